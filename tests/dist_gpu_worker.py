"""torchrun worker (N GPUs, NCCL): ONE sample sharded over N ranks (itsxpress_b200.distributed.run_sharded on
libitsx_b200, csrc/shard.cu) must equal the one-GPU path on the same sample -- global representatives, strands, trim
bounds and the re-expanded bytes of every block.  Prints one SHARDED_PARITY line per configuration.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
      tests/dist_gpu_worker.py [config:scale ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import synth
    from itsxpress_b200 import _lib
    from itsxpress_b200.distributed import Comm, GpuEngine, block_range, run_sharded
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    local_ctx, owner_ctx, single = _lib.Context(local), _lib.Context(local), _lib.Context(local)
    all_ok = True
    for spec in (sys.argv[1:] or ["c2_small:1.0", "c4s:0.05", "c3s:0.05"]):
        name, scale = spec.split(":")
        seq, off, which, cfg = synth.make_config(name, scale=float(scale))
        # reverse-complement copies of some reads at the end, so that classes with both strands span ranks
        n0 = len(off) - 1
        comp = np.zeros(256, np.uint8)
        comp[:] = np.arange(256)
        for a, b in zip(b"ACGT", b"TGCA"):
            comp[a] = b
        extra = [comp[seq[off[i]:off[i + 1]]][::-1] for i in range(0, min(n0, 400), 7)]
        seq = np.concatenate([seq] + extra)
        off = np.concatenate([off, off[-1] + np.cumsum([len(e) for e in extra])]).astype(np.int64)
        qual = synth.make_quals(9, off)
        n = len(off) - 1
        paths = [os.path.join(synth.HMM_DIR, f) for f in cfg["search_files"]]
        pre = [cfg["left_prefix"], cfg["right_prefix"]]
        for c in (owner_ctx, single):
            c.load_profiles(paths, pre)
            c.set_sides_by_prefix(*pre)
        lo, hi = block_range(n, rank, world)
        b0, b1 = int(off[lo]), int(off[hi])
        eng = GpuEngine(local_ctx, owner_ctx)
        got = run_sharded(eng, Comm(), seq[b0:b1], off[lo:hi + 1] - b0, lo)
        eng.upload(seq[b0:b1], off[lo:hi + 1] - b0, qual[b0:b1])
        gat = run_sharded(eng, Comm(), None, None, lo, want_rep=False, gather=True)
        want, st = single.run(seq, off)             # every rank also runs the whole sample alone
        want = {k: v.copy() for k, v in want.items()}
        _, wstrand, _ = single.derep(seq, off)
        wt, st2 = single.run_trim(seq, qual, off)
        a, b = np.searchsorted(wt["kept_index"], lo), np.searchsorted(wt["kept_index"], hi)
        o = wt["out_off"]
        ok = (np.array_equal(got["rep"], want["rep"][lo:hi]) and np.array_equal(got["keep"], want["keep"][lo:hi]) and
              np.array_equal(got["lo"], want["lo"][lo:hi]) and np.array_equal(got["hi"], want["hi"][lo:hi]) and
              np.array_equal(got["strand"], wstrand[lo:hi]) and got["n_unique_global"] == st.n_unique and
              np.array_equal(gat["kept_index"] + lo, wt["kept_index"][a:b]) and
              np.array_equal(gat["out_seq"], wt["out_seq"][o[a]:o[b]]) and
              np.array_equal(gat["out_qual"], wt["out_qual"][o[a]:o[b]]) and
              np.array_equal(gat["out_off"], o[a:b + 1] - o[a]))
        t = torch.tensor([1 if ok else 0, int(got["n_owned"]), int(wstrand[lo:hi].sum())], device="cuda")
        tmin = t.clone()
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        good = int(tmin[0].item()) == 1 and int(t[1].item()) == st.n_unique
        all_ok = all_ok and good
        if rank == 0:
            print("SHARDED_PARITY %s world=%d config=%s scale=%s reads=%d uniques=%d kept=%d trimmed_bytes=%d "
                  "minus_strand_reads=%d profiles=%d" % ("OK" if good else "FAIL", world, name, scale, n, st.n_unique,
                                                         st.n_kept, st2.out_bytes, int(t[2].item()), len(owner_ctx.names)),
                  flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()
