"""itsx_gzip_compress (deflate.cu) on the GPU: streams inflate to the input, are byte-identical to the CPU emulation of
the same per-thread code, and the writers use them."""
import gzip
import os

import numpy as np
import pytest

from test_gzip_core import build_emulator, emulate, gzip_cases

pytestmark = pytest.mark.gpu


def test_gpu_streams_inflate_and_equal_the_emulation(gpu_ctx, tmp_path):
    exe = build_emulator(tmp_path)
    for name, data in gzip_cases():
        raw = gpu_ctx.gzip_compress(data).tobytes()
        assert gzip.decompress(raw) == data, name
        assert raw == emulate(exe, tmp_path, data), name


def test_gpu_gzip_many_members_and_device_source(gpu_ctx):
    """more members than one launch takes (2 048), and a source that already lives in HBM"""
    import synth
    seq, off, which, cfg = synth.make_config("c2", scale=0.3)
    qual = synth.make_quals(3, off)
    text = np.concatenate([seq, qual, seq[::-1], qual[::-1]])[:72_000_000]
    raw = gpu_ctx.gzip_compress(text)
    assert len(raw) < 0.7 * text.size
    assert gzip.decompress(raw.tobytes()) == text.tobytes()
    import torch
    dev = torch.from_numpy(text[:5_000_000].copy()).cuda()
    from itsxpress_b200 import _lib
    L = _lib.lib()
    import ctypes as C
    cap = int(L.itsx_gzip_bound(dev.numel()))
    out = np.empty(cap, np.uint8)
    n = C.c_int64()
    torch.cuda.synchronize()
    gpu_ctx._chk(L.itsx_gzip_compress(gpu_ctx._h, C.c_void_p(dev.data_ptr()), dev.numel(), out.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
    assert gzip.decompress(out[:n.value].tobytes()) == text[:5_000_000].tobytes()


def test_writers_compress_on_the_gpu(tmp_path):
    from itsxpress_b200 import fastq as fq
    assert fq.GZIP_BACKEND == "gpu"
    with open(os.path.join(os.path.dirname(__file__), "test_data", "4774-1-MSITS3_merged.fastq"), "rb") as f:
        raw = f.read()
    fq.write_compressed(str(tmp_path / "a.gz"), raw, gzipped=True)
    assert gzip.open(str(tmp_path / "a.gz"), "rb").read() == raw
    w = fq.ChunkWriter(str(tmp_path / "b.gz"), gzipped=True)
    for i in range(0, len(raw), 50_000):
        w.write(raw[i:i + 50_000])
    w.close()
    assert gzip.open(str(tmp_path / "b.gz"), "rb").read() == raw
    w = fq.ChunkWriter(str(tmp_path / "e.gz"), gzipped=True)
    w.close()
    assert gzip.open(str(tmp_path / "e.gz"), "rb").read() == b""
    b = fq.read_fastq(str(tmp_path / "a.gz"))
    assert b.n == 227
