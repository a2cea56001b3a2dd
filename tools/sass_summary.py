#!/usr/bin/env python
"""Static facts about the kernels of the built library, no GPU needed: registers / spills / shared memory per kernel from
the `-Xptxas -v` logs of the build (itsxpress_b200/csrc/build/*.log), and a histogram of the SASS mnemonics that matter
(`cuobjdump -sass`) for the hot kernels.  python tools/sass_summary.py > profiles/<name>.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "itsxpress_b200", "csrc", "build")
LIB = os.path.join(ROOT, "itsxpress_b200", "libitsx_b200.so")
HOT = ["msv_kernel", "vit_kernel", "bias_kernel", "fb_kernel", "fbdec_kernel", "env_kernel", "mdfwd_kernel", "mdtrace_kernel",
       "mdclust_kernel", "merge_kernel", "deflate_kernel", "gather_kernel", "hash_kernel", "insert_kernel", "verify_kernel",
       "pack2_kernel"]
WATCH = ["VIADDMNMX", "VIMNMX3", "VIMNMX", "FFMA", "FMUL", "FADD", "FMNMX", "MUFU", "LDCU", "LDC", "LDG", "STG", "LDS", "STS", "LDL", "STL",
         "UBLKCP", "SYNCS", "ATOMG", "ATOMS", "RED", "SHFL", "BAR", "IMAD", "IADD3", "LOP3", "SHF", "PRMT", "POPC", "BRA"]


_NAMES = {}


def demangle(name):
    """kernel base name of a mangled symbol (cu++filt; template arguments kept, namespaces and parameters dropped)"""
    if name not in _NAMES:
        full = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        head = re.sub(r"^(?:void|int)\s+", "", full)
        depth, cut = 0, len(head)
        for i, c in enumerate(head):                 # the parameter list starts at the first '(' outside '<...>'
            if c == "<":
                depth += 1
            elif c == ">":
                depth -= 1
            elif c == "(" and depth == 0 and not head.startswith("<unnamed>", max(0, i - 9)):
                cut = i
                break
        head = head[:cut]
        _NAMES[name] = head.replace("<unnamed>::", "").strip()
    return _NAMES[name]


def ptxas():
    rows = []
    for log in sorted(os.listdir(BUILD)):
        if not log.endswith(".cu.log"):
            continue
        text = open(os.path.join(BUILD, log)).read()
        for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                             r"ptxas info\s+: Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", text):
            if "_kernel" not in m.group(1) or "cub" in m.group(1)[:40]:
                continue
            rows.append((log[:-4], demangle(m.group(1)), int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(8) or 0)))
    return rows


def sass():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    hist, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = demangle(m.group(1))
            hist.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Za-z0-9_]+)*)", line)
        if m and cur:
            hist[cur][m.group(1)] += 1
            hist[cur]["__all__"] += 1
            if ".S16x2" in m.group(2):
                hist[cur]["*.S16x2"] += 1
            if m.group(1) == "LDCU" and ".128" in m.group(2):
                hist[cur]["LDCU.128"] += 1
    return hist


def main():
    print("# Static kernel facts of the built library (no GPU involved)\n")
    print("`tools/sass_summary.py`: `-Xptxas -v` of the in-tree build (`-gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false`) and")
    print("`cuobjdump -sass itsxpress_b200/libitsx_b200.so`.  Counts are static instructions in the kernel's SASS, not executed ones.\n")
    print("## Registers, spills, shared memory (ptxas)\n")
    print("| source | kernel | registers | stack B | spill st B | spill ld B | static smem B |")
    print("|---|---|---:|---:|---:|---:|---:|")
    for r in ptxas():
        print("| `%s` | `%s` | %d | %d | %d | %d | %d |" % r)
    h = sass()
    print("\n## SASS mnemonics of the hot kernels\n")
    cols = ["__all__"] + WATCH + ["*.S16x2", "LDCU.128"]
    print("| kernel | " + " | ".join("all" if c == "__all__" else c for c in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for k in HOT:
        for name in sorted(h):
            if name == k or name.startswith(k):
                print("| `%s` | " % name + " | ".join(str(h[name].get(c, 0)) for c in cols) + " |")
    tc = sum(v for name in h for m_, v in h[name].items() if m_.startswith(("UTCMMA", "TCGEN", "HMMA", "IMMA", "UTCHMMA")))
    print("\nTensor-core instructions in the library: %d (the path is a max-plus / sum-product recurrence and byte work: none expected)." % tc)


if __name__ == "__main__":
    sys.exit(main())
