#!/usr/bin/env python
"""bench.py -- headline benchmark of the ITSxpress hot path on B200.

One "step" = one pass of the whole hot path (exact derep -> profile-HMM cascade -> boundary selection
-> trim bounds for every read) over one synthetic sample of BASELINE.json configs[1]
(1 M single-end 250 bp ITS1 reads, 30 % unique).  F.hmm (Fungi) is missing from the reference mount,
so the largest present analogue M.hmm (Metazoa, 98 ITS1 profiles) stands in -- stated in `config`.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--scale S] [--impl reference]

value  reads/s with the reads resident in HBM when the timed region starts (itsx_run_resident)
e2e    reads/s through the C-ABI call with pinned HOST buffers (itsx_run: H2D + kernels + D2H)
N > 1  (torchrun, one rank per GPU): every rank processes its own sample (QIIME2 artifacts are
       many independent samples) -> weak scaling, no data-path collective; time = max over ranks.
--impl reference  times the CPU oracle (restatement of vsearch + hmmsearch + ItsPosition + trim, all host
       threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import synth  # noqa: E402

INT_OPS_PER_2CELLS = 3.0      # VIMNMX (xB), VIADDMNMX (score, floor), half a VIMNMX3 x2 (row max): one s16x2 MSV cell pair
FP_OPS_PER_CELL = 10.0        # Forward 11 + Backward 9 FMA-pipe instructions per M/I/D cell, averaged
ENV_OPS_PER_CELL = 11.0       # envelope: Forward 11 + Backward 9 + decoding 2 per row cell, counted as 2 cells
# issue rates measured on this pool's B200 with tools/ubench/ffma2.cu (profiles/r1_ubench_issue_rates.txt):
# warp-instructions per clock per SM at 1965 MHz
MEASURED_FFMA_PER_CLK_SM = 3.820
MEASURED_S16X2_PER_CLK_SM = 1.959


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            d.update(hbm_gbs=float(j["hbm_gbs"]), sm_max_mhz=float(j.get("sm_max_mhz", 1965.0)), source="measured")
        except Exception:
            pass
    f = d["sm_max_mhz"] * 1e6
    # cells/s = SMs x warp-instr/clk/SM x 32 lanes x f / instructions per cell
    d["int_gcups"] = 148 * MEASURED_S16X2_PER_CLK_SM * 32 * f * 2.0 / INT_OPS_PER_2CELLS / 1e9
    d["fp_gcups"] = 148 * MEASURED_FFMA_PER_CLK_SM * 32 * f / FP_OPS_PER_CELL / 1e9
    d["env_gcups"] = 148 * MEASURED_FFMA_PER_CLK_SM * 32 * f / ENV_OPS_PER_CELL / 1e9
    return d


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(seq, off, which, frac_mod=12, pick=7):
    """Bounded sample that keeps the workload's unique fraction: all reads of every 12th unique
    (~80 k reads, ~10 s on 16 cores)."""
    sel = np.flatnonzero((which % frac_mod) == pick)
    lens = (off[1:] - off[:-1])[sel]
    o = np.zeros(len(sel) + 1, np.int64)
    o[1:] = np.cumsum(lens)
    delta = np.repeat(off[sel] - o[:-1], lens)
    s = seq[delta + np.arange(int(o[-1]), dtype=np.int64)]
    return s, o, "all reads of every %dth unique (%d reads, %d uniques)" % (frac_mod, len(sel),
                                                                             len(np.unique(which[sel])))


def oracle_pipeline(O, db, side, seq, off, threads=0):
    """derep -> search -> ItsPosition -> trim bounds on the CPU oracle; returns kept count."""
    rep, strand, nu = O.derep(seq, off)
    idx = np.flatnonzero(rep == np.arange(len(rep)))
    lens = (off[1:] - off[:-1])[idx]
    uoff = np.zeros(len(idx) + 1, np.int64)
    uoff[1:] = np.cumsum(lens)
    delta = np.repeat(off[idx] - uoff[:-1], lens)
    useq = seq[delta + np.arange(int(uoff[-1]), dtype=np.int64)]
    codes = O.digitize(useq.tobytes())
    rows, nrep, st = db.search(codes, uoff, O.default_params(threads))
    pos = O.itspos(rows, side, lens.astype(np.int32))
    n = len(rep)
    s_r = np.full(n, -1, np.int32); e_r = s_r.copy(); t_r = s_r.copy()
    s_r[idx] = pos["start"]; e_r[idx] = pos["stop"]; t_r[idx] = pos["tlen"]
    keep, lo, hi = O.trim_bounds(off, rep, s_r, e_r, t_r, mode=0)
    return int(keep.sum()), st


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle as O
    O.lib()
    seq, off, which, cfg = synth.make_config(args.config, scale=args.scale)
    s, o, desc = cpu_sample(seq, off, which)
    paths = [os.path.join(synth.HMM_DIR, f) for f in cfg["search_files"]]
    db = O.ProfileDB(paths, [cfg["left_prefix"], cfg["right_prefix"]])
    side = np.array([0 if n.startswith(cfg["left_prefix"]) else 1 for n in db.names], np.int8)
    cores = os.cpu_count()       # explicit: torchrun exports OMP_NUM_THREADS=1
    for _ in range(min(args.warmup, 1)):
        oracle_pipeline(O, db, side, s, o, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_pipeline(O, db, side, s, o, threads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = (len(o) - 1) / dt
    line = {"impl": "reference", "metric": "reads/s", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
            "config": workload_config(cfg, args),
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(cfg, args):
    length = cfg["length"]
    lmean = sum(length) / 2 if isinstance(length, tuple) else length
    return {"workload": "%s: %d reads of %s bp, %d unique (Zipf s=1), --region %s, profiles = %s %s_/%s_ "
                        "(F.hmm missing from the reference mount)" %
                        ({"c2": "BASELINE configs[1]", "c2_small": "configs[1] reduced", "c4s": "configs[3] shape, scaled",
                          "c3s": "configs[2] shape, scaled"}[args.config], cfg["n_reads"], length, cfg["n_unique"],
                         cfg["region"], "+".join(cfg["search_files"]) if len(cfg["search_files"]) < 4 else
                         "%d taxon files" % len(cfg["search_files"]), cfg["left_prefix"][0], cfg["right_prefix"][0]),
            "taxa": cfg["taxa"], "region": cfg["region"], "scale": args.scale,
            "l2_policy": "inputs (%.0f MB of read bytes per step) are larger than the 126 MB L2" %
                         (cfg["n_reads"] * lmean / 1e6),
            "parallelism": "1 sample per GPU (independent samples, no data-path collective)"}


def run_sharded_bench(args, ctx, rank, world, local):
    """One sample, block-partitioned over the ranks: itsxpress_b200.distributed.run_sharded (host buffers in and
    out on every rank, three collectives).  Timed by wall clock between barriers, max over ranks."""
    import torch
    import torch.distributed as dist
    from itsxpress_b200 import _lib
    from itsxpress_b200.distributed import block_range, run_sharded_device
    seq, off, which, cfg = synth.make_config(args.config, scale=args.scale)
    n = len(off) - 1
    ctx.load_profiles([os.path.join(synth.HMM_DIR, f) for f in cfg["search_files"]], [cfg["left_prefix"], cfg["right_prefix"]])
    ctx.set_sides_by_prefix(cfg["left_prefix"], cfg["right_prefix"])
    lo, hi = block_range(n, rank, world)
    bseq = np.ascontiguousarray(seq[off[lo]:off[hi]])
    boff = off[lo:hi + 1] - off[lo]
    prm = _lib.default_params()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    for _ in range(max(args.warmup, 3)):
        out = run_sharded_device(ctx, bseq, boff, lo, prm)
    clk = ClockSampler(local)
    clk.start()
    l0 = ctx.launch_count()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = run_sharded_device(ctx, bseq, boff, lo, prm)
    barrier()
    dt = time.perf_counter() - t0
    launches = ctx.launch_count() - l0
    clocks = clk.stop()
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    k = torch.tensor([int(out["keep"].sum())], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(k, op=dist.ReduceOp.SUM)
    dt = float(t.item())
    if rank == 0:
        v = n / (dt / args.steps)
        cfgd = workload_config(cfg, args)
        cfgd["parallelism"] = ("one sample sharded over %d GPU(s), device resident: local derep -> all-to-all of local uniques to "
                               "owner key%%G -> owner derep + HMM search -> all-reduce domZ -> all-gather positions "
                               "-> local trim" % world)
        line = {"metric": "reads/s", "value": v, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u8/s16x2 (MSV) + f32 (Forward/Backward)",
                "data": "synthetic", "config": cfgd, "clocks": clocks,
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": int(seq.nbytes + off.nbytes),
                        "d2h_bytes_per_step": int(n * 13), "ms_per_step": dt / args.steps * 1e3},
                "gpu_launches": int(launches), "roofline": None, "cpu_baseline": None,
                "result": {"n_unique": int(out["n_unique_global"]), "n_kept": int(k.item())},
                "note": "sharded mode: value == e2e (wall clock between barriers; the rank's block of reads goes in from "
                        "host memory, keep/lo/hi/rep come back to it; every intermediate and the three exchanges stay on "
                        "the devices, NCCL over NVLink)"}
        emit(line)


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything else that libraries print (NCCL's version
    banner, torch warnings) was moved to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (testing only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c2", choices=["c2", "c4s", "c3s", "c2_small"],
                    help="workload: c2 = BASELINE configs[1] (the headline); c4s / c3s = scaled shapes of configs[3] / [2]")
    ap.add_argument("--sharded", action="store_true",
                    help="ONE sample sharded over all ranks (hash-partitioned derep all-to-all, domZ all-reduce, "
                         "position all-gather; strong scaling) instead of one sample per rank")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from itsxpress_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.Context(local)
    if args.sharded:
        run_sharded_bench(args, ctx, rank, world, local)
        if world > 1:
            dist.destroy_process_group()
        return
    # every rank gets its own sample (different seed)
    seq, off, which, cfg = synth.make_config(args.config, seed=2 * 1_000_003 + rank, scale=args.scale)
    nreads = len(off) - 1
    ctx.load_profiles([os.path.join(synth.HMM_DIR, f) for f in cfg["search_files"]], [cfg["left_prefix"], cfg["right_prefix"]])
    ctx.set_sides_by_prefix(cfg["left_prefix"], cfg["right_prefix"])
    prm = _lib.default_params()

    # pinned host staging for the e2e leg
    pin_in = _lib.PinnedBuffer(seq.nbytes + off.nbytes + 64)
    pseq = pin_in.array(np.uint8, seq.size)
    poff = pin_in.array(np.int64, off.size, offset=(seq.nbytes + 63) // 64 * 64)
    pseq[:] = seq
    poff[:] = off
    pin_out = _lib.PinnedBuffer(nreads * 13 + 256)
    o_rep = pin_out.array(np.int32, nreads, 0)
    o_lo = pin_out.array(np.int32, nreads, nreads * 4)
    o_hi = pin_out.array(np.int32, nreads, nreads * 8)
    o_keep = pin_out.array(np.uint8, nreads, nreads * 12)
    outs = dict(rep=o_rep, keep=o_keep, lo=o_lo, hi=o_hi)

    ext = torch.cuda.ExternalStream(ctx.stream, device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        with torch.cuda.stream(ext):
            e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        barrier()
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    # ---- resident leg (value) ----
    ctx.reads_upload(pseq, poff)
    stage = {k: 0.0 for k in ("ms_msv", "ms_bias", "ms_fwd", "ms_mdom", "ms_env", "ms_final", "ms_total")}
    dstage = {k: 0.0 for k in ("ms_pack", "ms_hash", "ms_insert", "ms_verify", "ms_compact", "ms_total")}
    last = {}

    def step_resident():
        st = ctx.run_resident(prm)
        ss, ds = ctx.search_stats(), ctx.derep_stats()
        for k in stage:
            stage[k] += getattr(ss, k)
        for k in dstage:
            dstage[k] += getattr(ds, k)
        last["run"], last["search"], last["derep"] = st, ss, ds

    for _ in range(max(args.warmup, 3)):
        step_resident()
    for k in stage:
        stage[k] = 0.0
    for k in dstage:
        dstage[k] = 0.0
    clk = ClockSampler(local)
    clk.start()
    l0 = ctx.launch_count()
    ms_dev, ms_wall = timed(step_resident, args.steps)
    launches = ctx.launch_count() - l0
    # ---- e2e leg: host buffers in, host buffers out ----
    for _ in range(2):
        ctx.run(pseq, poff, prm, out=outs)
    ms_e2e_dev, ms_e2e_wall = timed(lambda: ctx.run(pseq, poff, prm, out=outs), args.steps)
    clocks = clk.stop()

    total_reads = nreads * world
    value = total_reads / (ms_dev / 1e3 / args.steps)
    e2e_v = total_reads / (max(ms_e2e_dev, ms_e2e_wall) / 1e3 / args.steps)
    ss, ds, rs = last["search"], last["derep"], last["run"]
    pk = peaks()
    K = args.steps
    sec = {k: v / 1e3 / K for k, v in stage.items()}
    dsec = {k: v / 1e3 / K for k, v in dstage.items()}
    L = sum(cfg["length"]) / 2 if isinstance(cfg["length"], tuple) else cfg["length"]
    stages = {
        "msv": {"ms": sec["ms_msv"] * 1e3, "gcups": ss.msv_cells / max(sec["ms_msv"], 1e-9) / 1e9,
                "peak_gcups": pk["int_gcups"], "bound": "int-alu (s16x2 DPX)"},
        "fwd_bwd_decode": {"ms": sec["ms_fwd"] * 1e3,
                           "gcups": (ss.fwd_cells + ss.bck_cells) / max(sec["ms_fwd"], 1e-9) / 1e9,
                           "peak_gcups": pk["fp_gcups"], "bound": "fp32 fma"},
        "envelope": {"ms": sec["ms_env"] * 1e3, "gcups": ss.env_cells / max(sec["ms_env"], 1e-9) / 1e9,
                     "peak_gcups": pk["env_gcups"], "bound": "fp32 fma (+ TMA-staged scratch rows)"},
        "bias": {"ms": sec["ms_bias"] * 1e3, "rows_per_s": ss.bias_rows / max(sec["ms_bias"], 1e-9)},
        "multidomain": {"ms": sec["ms_mdom"] * 1e3, "regions": ss.n_multidomain_regions,
                        "traces_per_s": ss.n_multidomain_regions * 200 / max(sec["ms_mdom"], 1e-9),
                        "bound": "latency / issue (200 stochastic tracebacks per flagged region, p7_domaindef)"},
        "derep_pack": {"ms": dsec["ms_pack"] * 1e3,
                       "gbs": ds.bytes_ascii * (1 + 0.25 + 0.125) / max(dsec["ms_pack"], 1e-9) / 1e9,
                       "peak_gbs": pk["hbm_gbs"], "bound": "hbm"},
        "derep_hash": {"ms": dsec["ms_hash"] * 1e3,
                       "gbs": (ds.bytes_ascii * 0.25 + 25.0 * nreads) / max(dsec["ms_hash"], 1e-9) / 1e9,
                       "peak_gbs": pk["hbm_gbs"], "bound": "hbm"},
        "derep_insert_verify": {"ms": (dsec["ms_insert"] + dsec["ms_verify"]) * 1e3,
                                "gbs": nreads * (L / 2 + 28.0) / max(dsec["ms_insert"] + dsec["ms_verify"], 1e-9) / 1e9,
                                "peak_gbs": pk["hbm_gbs"], "bound": "hbm + atomics"},
        "search_total_ms": sec["ms_total"] * 1e3, "derep_total_ms": dsec["ms_total"] * 1e3,
        "survival": {"pairs": ss.n_pairs, "past_msv": ss.n_past_msv, "past_bias": ss.n_past_bias,
                     "past_fwd": ss.n_past_fwd, "hits": ss.n_hits_reported, "domains": ss.n_domains,
                     "multidomain_regions": ss.n_multidomain_regions},
    }
    dom = max(("msv", "fwd_bwd_decode", "envelope"), key=lambda k: stages[k]["ms"])
    # dram traffic of the dominant kernel from the committed ncu --set full capture (one per-profile launch at
    # --scale 0.2; profiles/*_fb_kernel_full.md): far below anything HBM-bound -- the DP kernels are issue-bound
    traffic = {"fb_kernel": 262.9e6, "env_kernel": 376.1e6, "msv_kernel": 9.0e6}       # profiles/r1e_*_full.md
    kname = {"msv": "msv_kernel", "fwd_bwd_decode": "fb_kernel", "envelope": "env_kernel"}[dom]
    roof = {"kernel": kname, "bound": "alu", "achieved": stages[dom]["gcups"], "peak": stages[dom]["peak_gcups"],
            "unit": "GCUPS", "frac": stages[dom]["gcups"] / stages[dom]["peak_gcups"], "traffic": traffic[kname],
            "traffic_note": "dram bytes of one launch in the ncu capture at --scale 0.2 (per-profile launch)",
            "note": "DP recurrence: fp32 / s16x2 issue-bound, not HBM or tensor; peak = MEASURED issue rate "
                    "(tools/ubench/ffma2.cu: %.3f FFMA, %.3f VIADDMNMX.S16x2 warp-instr/clk/SM) x 148 SMs x "
                    "clocks.max.sm / instructions per cell; %s HBM peak %.0f GB/s is used for the derep kernels in "
                    "`stages`" % (MEASURED_FFMA_PER_CLK_SM, MEASURED_S16X2_PER_CLK_SM, pk["source"], pk["hbm_gbs"])}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as O
        s, o, desc = cpu_sample(seq, off, which)
        db = O.ProfileDB([os.path.join(synth.HMM_DIR, f) for f in cfg["search_files"]], [cfg["left_prefix"], cfg["right_prefix"]])
        side = np.array([0 if n.startswith(cfg["left_prefix"]) else 1 for n in db.names], np.int8)
        t0 = time.perf_counter()
        kept, ost = oracle_pipeline(O, db, side, s, o, threads=os.cpu_count())
        dt = time.perf_counter() - t0
        cpu = {"value": (len(o) - 1) / dt, "unit": "reads/s", "cores": os.cpu_count(), "kind": "port",
               "sample": desc, "seconds": dt}

    if rank == 0:
        line = {"metric": "reads/s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/s16x2 (MSV) + f32 (Forward/Backward)",
                "data": "synthetic", "config": workload_config(cfg, args), "clocks": clocks,
                "e2e": {"value": e2e_v, "unit": "reads/s", "h2d_bytes_per_step": int(seq.nbytes + off.nbytes),
                        "d2h_bytes_per_step": int(nreads * 13), "ms_per_step": max(ms_e2e_dev, ms_e2e_wall) / args.steps},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "stages": stages,
                "hmm_gcups": {"msv": stages["msv"]["gcups"], "fwd_bwd": stages["fwd_bwd_decode"]["gcups"],
                              "envelope": stages["envelope"]["gcups"]},
                "result": {"n_unique": int(rs.n_unique), "n_kept": int(rs.n_kept)}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
