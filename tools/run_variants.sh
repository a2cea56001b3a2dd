#!/bin/bash
# Bench several builds of the library back to back on one GPU (experiments; see tools/build_variant.py).
#   tools/run_variants.sh TAG base pf8 ring8 ...     ("base" = itsxpress_b200/libitsx_b200.so)
# Lines go to gpurun_out/TAG_<variant>.json; a summary of stage ms and result checksums is printed at the end.
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset ITSX_B200_LIB; else export ITSX_B200_LIB=$PWD/tools/variants/$v.so; fi
  timeout 300 python bench.py --steps ${STEPS:-3} --warmup 3 --no-cpu-baseline --no-cli ${BENCH_ARGS} \
      > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err || echo "variant $v failed: $(tail -3 gpurun_out/${TAG}_$v.err)"
done
python - "$TAG" "$@" <<'PY'
import json, sys
tag, vs = sys.argv[1], sys.argv[2:]
for v in vs:
    try:
        d = json.loads(open("gpurun_out/%s_%s.json" % (tag, v)).read().strip().splitlines()[-1])
    except Exception as e:
        print(v, "no line", e); continue
    s = d["stages"]
    print("%-8s step %.1f ms | msv %.1f bias %.1f fb %.1f md %.1f env %.1f | kept %d crc %s msel %s" % (
        v, d["ms_per_step"], s["msv"]["ms"], s["bias"]["ms"], s["fwd_bwd_decode"]["ms"], s["multidomain"]["ms"],
        s["envelope"]["ms"], d["result"]["n_kept"], d["result"].get("crc32"), d["result"].get("n_selected_multidomain")))
PY
