import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TD = os.path.join(ROOT, "tests", "test_data")
HMM_DIR = os.path.join(ROOT, "itsxpress_b200", "ITSx_db", "HMMs")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")
    # the CPU suite (-m "not gpu", or a box without a device) exercises the host side of the writers with zlib members;
    # the GPU suite leaves the default: .gz output is compressed by itsx_gzip_compress (tests/test_gpu_gzip.py)
    m = config.getoption("-m") or ""
    if "not gpu" in m or (m.strip() != "gpu" and not _have_gpu()):
        os.environ["ITSX_GZIP"] = "host"


_HAVE_GPU = None


def _have_gpu():
    """True iff libitsx_b200 can open device 0 (a B200); decided once per session."""
    global _HAVE_GPU
    if _HAVE_GPU is None:
        try:
            from itsxpress_b200 import _lib
            _lib.Context(0).close()
            _HAVE_GPU = True
        except Exception:
            _HAVE_GPU = False
    return _HAVE_GPU


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a CPU box skips the gpu-marked tests instead of erroring in every one of them.  With
    `-m gpu` (the GPU box) nothing is skipped: a missing device must fail loudly there."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if gpu_items and not _have_gpu():
        skip = pytest.mark.skip(reason="no B200 in this box (gpu-marked tests run with -m gpu on the GPU box)")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def gpu_ctx():
    from itsxpress_b200 import _lib
    ctx = _lib.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def fixture_reads():
    from itsxpress_b200.fastq import read_fastq
    b = read_fastq(os.path.join(TD, "ex_tmpdir", "seq.fq.gz"))
    seq, off = b.seq_concat()
    qual, _ = b.qual_concat()
    return b, seq, off, qual
