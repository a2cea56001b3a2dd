#!/usr/bin/env python
"""BASELINE configs[4]: QIIME 2 trim-pair-output-unmerged on a 50 M-read-pair artifact (5 % unique within a sample),
all GPUs of the box.  SURVEY 8(d) fixes the shape the BASELINE leaves open: 64 samples x 781 250 pairs, 2 x 250 bp off
330-441 bp ITS2 amplicons, gzipped Casava files in, gzipped trimmed R1 + R2 files out.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/c5_artifact_bench.py \
      [--samples 64] [--pairs 781250] [--distinct-per-rank 1]

Every rank synthesises `--distinct-per-rank` samples (seeded by rank) and hard-links them under the other sample ids it
contributes (samples are processed independently -- derep, Z and domZ are per sample upstream, q2_itsxpress.py:273-296 --
so identical content under different ids costs exactly what distinct content would; it only shortens the untimed
synthesis).  Timed region: itsxpress_b200.q2_itsxpress.main_sharded over the whole artifact (samples dealt to ranks,
every rank: inflate -> merge -> derep -> search -> trim -> deflate), barrier to barrier, max over ranks.
Prints one JSON line on rank 0.
"""
import argparse
import gzip
import json
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth  # noqa: E402
from cli_e2e import write_fastq  # noqa: E402


def make_sample(seed, n, dst1, dst2):
    frag, foff, _, _ = synth.make_reads(seed, n, max(300, n // 20), (330, 441), "M.hmm", "3_", "4_", zipf_s=1.2,
                                        spacer=(150, 230))
    fs, fq_, fo, rs, rq, ro = synth.make_pairs(seed + 1, frag, foff, read_len=250, err_scale=0.0, n_rate=0.0)
    for dst, (s, o, q, mate) in ((dst1, (fs, fo, fq_, 1)), (dst2, (rs, ro, rq, 2))):
        plain = dst[:-3]
        write_fastq(plain, s, o, q, mate)
        with open(plain, "rb") as f, gzip.open(dst, "wb", compresslevel=1) as g:
            shutil.copyfileobj(f, g, 1 << 24)
        os.remove(plain)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--pairs", type=int, default=781_250)
    ap.add_argument("--distinct-per-rank", type=int, default=1)
    ap.add_argument("--dir", default="/tmp/itsx_c5")
    ap.add_argument("--gzip", default="gpu", help="who deflates the outputs, timed one after the other on the same inputs: "
                                                  "'gpu' (itsx_gzip_compress), 'host' (zlib members on the host cores) or 'host,gpu'")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    src, out = os.path.join(a.dir, "in"), os.path.join(a.dir, "out")
    if rank == 0:
        shutil.rmtree(a.dir, ignore_errors=True)
        os.makedirs(src)
    barrier()
    # ---- untimed: synthesis ----
    t0 = time.perf_counter()
    mine = [k for k in range(a.samples) if k % world == rank]
    made = []
    for j, k in enumerate(mine):
        f1 = os.path.join(src, "S%03d_%d_L001_R1_001.fastq.gz" % (k, k))
        f2 = os.path.join(src, "S%03d_%d_L001_R2_001.fastq.gz" % (k, k))
        if j < a.distinct_per_rank:
            make_sample(5 * 1_000_003 + 17 * k, a.pairs, f1, f2)
            made.append((f1, f2))
        else:
            g1, g2 = made[j % len(made)]
            os.link(g1, f1)
            os.link(g2, f2)
    t_synth = time.perf_counter() - t0
    barrier()
    if rank == 0:
        lines = ["sample-id,filename,direction"]
        for k in range(a.samples):
            lines.append("S%03d,S%03d_%d_L001_R1_001.fastq.gz,forward" % (k, k, k))
            lines.append("S%03d,S%03d_%d_L001_R2_001.fastq.gz,reverse" % (k, k, k))
        with open(os.path.join(src, "MANIFEST"), "w") as f:
            f.write("\n".join(lines) + "\n")
        with open(os.path.join(src, "metadata.yml"), "w") as f:
            f.write("{phred-offset: 33}\n")
    barrier()
    from itsxpress_b200 import q2_itsxpress as q2
    from itsxpress_b200 import SeqSample
    # warm-up: CUDA context, profile tables, host thread pools (one small sample through the same code path)
    warm = os.path.join(a.dir, "warm_r%d" % rank)
    os.makedirs(warm)
    wf1, wf2 = os.path.join(warm, "W_0_L001_R1_001.fastq.gz"), os.path.join(warm, "W_0_L001_R2_001.fastq.gz")
    make_sample(99 + rank, 20_000, wf1, wf2)
    with open(os.path.join(warm, "MANIFEST"), "w") as f:
        f.write("sample-id,filename,direction\nW,W_0_L001_R1_001.fastq.gz,forward\nW,W_0_L001_R2_001.fastq.gz,reverse\n")
    q2.main_sharded(q2.PerSampleDir(warm), os.path.join(warm, "out"), region="ITS2", taxa="M", rank=0, world=1,
                    barrier=lambda: None)
    ctx = SeqSample.get_context()
    from itsxpress_b200 import fastq as fqmod
    for backend in a.gzip.split(","):
        fqmod.GZIP_BACKEND = backend
        if rank == 0:
            shutil.rmtree(out, ignore_errors=True)
        barrier()
        l0 = ctx.launch_count()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, done = q2.main_sharded(q2.PerSampleDir(src), out, region="ITS2", taxa="M", paired_in=True, paired_out=True)
        torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        launches = ctx.launch_count() - l0
        t = torch.tensor([dt, float(launches), float(len(done))], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if rank == 0:
            dt = float(tmax[0].item())
            outs = [f for f in os.listdir(out) if f.endswith(".fastq.gz")]
            in_bytes = sum(os.path.getsize(os.path.join(src, f)) for f in os.listdir(src) if f.endswith(".gz"))
            out_bytes = sum(os.path.getsize(os.path.join(out, f)) for f in outs)
            n_out = 0
            with gzip.open(os.path.join(out, sorted(outs)[0]), "rb") as f:
                for _ in f:
                    n_out += 1
            print(json.dumps({
                "metric": "read pairs/s, QIIME 2 trim-pair-output-unmerged, artifact in -> artifact out",
                "value": a.samples * a.pairs / dt, "unit": "pairs/s", "n_gpus": world, "seconds": dt,
                "config": {"workload": "BASELINE configs[4]: %d samples x %d read pairs (2 x 250 bp off 330-441 bp ITS2 amplicons, "
                                       "5 %% unique within a sample), gzipped Casava files in and out, --region ITS2, profiles = "
                                       "M.hmm 3_/4_ (F.hmm missing from the reference mount)" % (a.samples, a.pairs),
                           "parallelism": "whole samples dealt to ranks (largest first), no data-path collective; per sample: "
                                          "inflate -> GPU merge -> GPU derep -> GPU search -> GPU trim -> deflate",
                           "distinct_samples": a.distinct_per_rank * world},
                "gzip_backend": backend, "host_cores": os.cpu_count(), "samples_done": int(t[2].item()), "gpu_launches": int(t[1].item()),
                "input_gz_bytes": in_bytes, "output_gz_bytes": out_bytes, "output_files": len(outs),
                "records_in_first_output": n_out // 4, "synthesis_seconds_rank0": t_synth,
                "bound": "host (device time per sample ~0.1 s): gzip inflate of the inputs" + (" and deflate of the outputs" if backend == "host" else ", FASTQ parsing / formatting")}))
        barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
