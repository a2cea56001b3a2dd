#!/usr/bin/env python
"""Build an experimental variant of libitsx_b200.so: search.cu recompiled with extra -D flags, the other objects reused.

  python tools/build_variant.py NAME -DFB_DECODE_PF=8 ...   ->  tools/variants/NAME.so

Run a bench or the tests against it with ITSX_B200_LIB=tools/variants/NAME.so (experiments only; the product
loads itsxpress_b200/libitsx_b200.so).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from itsxpress_b200 import build as B  # noqa: E402


def main():
    name, defs = sys.argv[1], sys.argv[2:]
    B.build()
    vdir = os.path.join(ROOT, "tools", "variants")
    os.makedirs(vdir, exist_ok=True)
    obj = os.path.join(vdir, name + ".search.o")
    cmd = [B.NVCC] + B.FLAGS + defs + ["-x", "cu", "-c", os.path.join(B.CSRC, "search.cu"), "-o", obj]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    open(os.path.join(vdir, name + ".log"), "w").write(out.stdout)
    if out.returncode:
        sys.stderr.write(out.stdout)
        raise SystemExit(1)
    objs = [os.path.join(B.CSRC, "build", s.rsplit(".", 1)[0] + ".o") for s in B.SOURCES if s != "search.cu"] + [obj]
    lib = os.path.join(vdir, name + ".so")
    subprocess.check_call([B.NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lpthread"])
    for line in out.stdout.splitlines():
        if "fb_kernel" in line or "env_kernel" in line or "mdclust" in line:
            print(line[:60])
    print(lib)


if __name__ == "__main__":
    main()
