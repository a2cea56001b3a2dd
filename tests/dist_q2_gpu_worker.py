"""torchrun worker (N GPUs): q2_itsxpress.main_sharded deals the samples of a paired artifact to the ranks (one GPU
each, no data-path collective); every output file must equal what the plain single-process action writes.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 \
      tests/dist_q2_gpu_worker.py <workdir>"""
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TD = os.path.join(ROOT, "tests", "test_data")


def main():
    work = sys.argv[1]
    import torch.distributed as dist
    from itsxpress_b200 import fastq as fq
    from itsxpress_b200 import q2_itsxpress as q2
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")           # barriers only: the data path has no collective
    src = os.path.join(work, "in")
    if rank == 0:
        # a 5-sample artifact: the reference's paired sample, whole and as prefixes of different size
        shutil.rmtree(work, ignore_errors=True)
        os.makedirs(src)
        lines = ["sample-id,filename,direction"]
        b = [open(os.path.join(TD, "4774-1-MSITS3_R%d.fastq" % m)).readlines() for m in (1, 2)]
        import gzip
        for k, npairs in enumerate([250, 60, 200, 120, 30]):
            for m, tag, d in ((0, "R1", "forward"), (1, "R2", "reverse")):
                fn = "S%d_%d_L001_%s_001.fastq.gz" % (k, k, tag)
                with gzip.open(os.path.join(src, fn), "wt") as f:
                    f.write("".join(b[m][:4 * npairs]))
                lines.append("S%d,%s,%s" % (k, fn, d))
        open(os.path.join(src, "MANIFEST"), "w").write("\n".join(lines) + "\n")
        open(os.path.join(src, "metadata.yml"), "w").write("{phred-offset: 33}\n")
    dist.barrier()
    t0 = time.perf_counter()
    res, mine = q2.main_sharded(q2.PerSampleDir(src), os.path.join(work, "out"), region="ITS2", taxa="M")
    dt = time.perf_counter() - t0
    print("rank %d of %d: samples %s in %.2f s on GPU %s" % (rank, world, mine, dt, os.environ.get("LOCAL_RANK")), flush=True)
    dist.barrier()
    if rank == 0:
        ref = q2.trim_pair_output_unmerged(q2.PerSampleDir(src), region="ITS2", taxa="M")
        names = sorted(f for f in os.listdir(str(ref)) if f.endswith(".fastq.gz"))
        assert names == sorted(f for f in os.listdir(str(res)) if f.endswith(".fastq.gz")) and len(names) == 10
        for n in names:
            assert fq._open_bytes(os.path.join(str(res), n)) == fq._open_bytes(os.path.join(str(ref), n)), n
        assert open(os.path.join(str(res), "MANIFEST")).read() == open(os.path.join(str(ref), "MANIFEST")).read()
        print("q2 main_sharded over %d GPUs == single process: OK (%d files)" % (world, len(names)), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
