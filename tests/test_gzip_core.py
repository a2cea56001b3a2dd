"""The per-thread bodies of the GPU gzip writer (itsxpress_b200/csrc/deflate_core.h) on the CPU: tools/deflate_emul.cpp
runs the kernel's phases thread by thread; every stream must inflate (zlib checks CRC-32 and ISIZE of every member) to
the input.  The kernel itself is held to the same streams, byte for byte, in tests/test_gpu_gzip.py."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from itsxpress_b200 import fastq as fq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TD = os.path.join(ROOT, "tests", "test_data")


def build_emulator(tmp):
    exe = os.path.join(str(tmp), "deflate_emul")
    subprocess.check_call(["g++", "-O2", "-o", exe, os.path.join(ROOT, "tools", "deflate_emul.cpp")])
    return exe


def inflates_everywhere(raw, data):
    """zlib AND this repo's own reader (one core; several cores with the chunking shrunk to the input's size)"""
    return (gzip.decompress(raw) == data and fq.gunzip(raw, 1).tobytes() == data and
            fq.gunzip(raw, 4, (512, 2048, 0)).tobytes() == data)


def emulate(exe, tmp, data):
    src, dst = os.path.join(str(tmp), "in.bin"), os.path.join(str(tmp), "out.gz")
    with open(src, "wb") as f:
        f.write(data)
    subprocess.check_call([exe, src, dst], stderr=subprocess.DEVNULL)
    with open(dst, "rb") as f:
        return f.read()


def gzip_cases():
    rng = np.random.default_rng(11)
    fib = [1, 1]
    while len(fib) < 24:
        fib.append(fib[-1] + fib[-2])
    skew = np.concatenate([np.full(f, i, np.uint8) for i, f in enumerate(fib)])
    rng.shuffle(skew)
    with open(os.path.join(TD, "4774-1-MSITS3_merged.fastq"), "rb") as f:
        fastq = f.read()
    return [
        ("empty", b""), ("one byte", b"A"), ("three bytes", b"ACG"), ("run", b"I" * 100000),
        ("one chunk", bytes(rng.integers(65, 69, 32768, dtype=np.uint8))),
        ("one chunk and a byte", bytes(rng.integers(65, 69, 32769, dtype=np.uint8))),
        ("incompressible", bytes(rng.integers(0, 256, 100000, dtype=np.uint8))),          # stored blocks
        ("two symbols", bytes(rng.integers(0, 2, 70000, dtype=np.uint8))),
        ("fibonacci frequencies", skew.tobytes()),                                         # code length limit
        ("all byte values", bytes(np.minimum(rng.geometric(0.02, 200000), 255).astype(np.uint8))),
        ("fastq", fastq),
    ]


def test_emulated_streams_inflate_to_the_input(tmp_path):
    exe = build_emulator(tmp_path)
    for name, data in gzip_cases():
        raw = emulate(exe, tmp_path, data)
        assert inflates_everywhere(raw, data), name
        assert raw[:4] == b"\x1f\x8b\x08\x00", name
    # amplicon FASTQ compresses about as well as zlib level 1
    name, data = gzip_cases()[-1]
    import zlib
    assert len(emulate(exe, tmp_path, data)) < 1.1 * len(zlib.compress(data, 1))


def test_emulated_streams_at_block_member_and_window_boundaries(tmp_path):
    """sizes around one block (32 768 bytes) and one member (32 blocks), repeats at distances around the 32 768-byte
    window and around the longest match (258), and mixtures of literals, runs and copies"""
    exe = build_emulator(tmp_path)
    rng = np.random.default_rng(5)
    C, G = 32768, 32
    cases = []
    for n in (C - 1, C + 1, 2 * C, G * C - 1, G * C, G * C + 1, (G + 1) * C + 5):
        cases.append(("acgt %d" % n, bytes(rng.integers(65, 69, n, dtype=np.uint8))))
    for dist in (1, 2, 3, 257, 258, 259, 32767, 32768, 32769):
        block = bytes(rng.integers(0, 256, dist, dtype=np.uint8))
        cases.append(("period %d" % dist, (block * (150000 // dist + 2))[:150000]))
    for it in range(12):
        out = bytearray()
        target, alpha = int(rng.integers(1, 4 * C)), int(rng.integers(2, 257))
        while len(out) < target:
            k = rng.integers(0, 3)
            if k == 0 or len(out) < 8:
                out += bytes(rng.integers(0, alpha, int(rng.integers(1, 300)), dtype=np.uint8))
            elif k == 1:
                out += bytes([int(rng.integers(0, alpha))]) * int(rng.integers(1, 700))
            else:
                dist, ln = int(rng.integers(1, min(len(out), 70000) + 1)), int(rng.integers(3, 600))
                for _ in range(ln):
                    out.append(out[-dist])
        cases.append(("mix %d" % it, bytes(out[:target])))
    for name, data in cases:
        assert inflates_everywhere(emulate(exe, tmp_path, data), data), name
