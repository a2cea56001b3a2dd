"""torchrun worker (N GPUs, NCCL): run_sharded with the GPU engine over N ranks must equal itsx_run on one GPU.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
      tests/dist_gpu_worker.py"""
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
    from itsxpress_b200.distributed import Comm, GpuEngine, block_range, run_sharded, run_sharded_device
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    seq, off, which, cfg = synth.make_config("c2_small", scale=1.0)
    ctx = _lib.Context(local)
    ctx.load_profiles([os.path.join(synth.HMM_DIR, cfg["hmm_file"])], [cfg["left_prefix"], cfg["right_prefix"]])
    ctx.set_sides_by_prefix(cfg["left_prefix"], cfg["right_prefix"])
    n = len(off) - 1
    lo, hi = block_range(n, rank, world)
    got = run_sharded(GpuEngine(ctx), Comm(), seq[off[lo]:off[hi]], off[lo:hi + 1] - off[lo], lo)
    dev = run_sharded_device(ctx, seq[off[lo]:off[hi]], off[lo:hi + 1] - off[lo], lo)   # device-resident exchange
    want, st = ctx.run(seq, off)             # every rank also runs the whole sample alone
    ok = True
    for res in (got, dev):
        ok = ok and (np.array_equal(res["rep"], want["rep"][lo:hi]) and np.array_equal(res["keep"], want["keep"][lo:hi]) and
                     np.array_equal(res["lo"], want["lo"][lo:hi]) and np.array_equal(res["hi"], want["hi"][lo:hi]) and
                     res["n_unique_global"] == st.n_unique)
    ok = ok and np.array_equal(got["strand"], dev["strand"]) and np.array_equal(got["nreported"], dev["nreported"])
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_PARITY %s world=%d reads=%d uniques=%d kept=%d" %
              ("OK" if int(t.item()) == 1 else "FAIL", world, n, st.n_unique, st.n_kept))
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
