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
