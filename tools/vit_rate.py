#!/usr/bin/env python
"""Rate of the Viterbi filter stage (K6) on the GPU: BASELINE configs[1]'s uniques under thresholds that send every
MSV + bias survivor through it (F1 = 0.3, F2 = 1e-4, F3 = 1e-5; the reference's F1 == F2 never runs the stage).
Prints one JSON line: pairs through the filter, DP cells, device ms, GCUPS.   python tools/vit_rate.py [--scale S]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.3)
    a = ap.parse_args()
    from itsxpress_b200 import _lib
    seq, off, which, cfg = synth.make_config("c2", scale=a.scale)
    ctx = _lib.Context(0)
    ctx.load_profiles([os.path.join(synth.HMM_DIR, cfg["hmm_file"])], [cfg["left_prefix"], cfg["right_prefix"]])
    ctx.set_sides_by_prefix(cfg["left_prefix"], cfg["right_prefix"])
    prm = _lib.default_params()
    prm.F1, prm.F2, prm.F3 = 0.3, 1e-4, 1e-5
    ctx.derep(seq, off)
    best = None
    for _ in range(3):
        ctx.search(prm)
        st = ctx.search_stats()
        if best is None or st.ms_vit < best.ms_vit:
            best = st
    print(json.dumps({"stage": "viterbi filter (vit_kernel)", "pairs_run": int(best.n_vit_run),
                      "pairs_past_bias": int(best.n_past_bias), "pairs_past_viterbi": int(best.n_past_vit),
                      "cells": best.vit_cells, "ms": best.ms_vit, "gcups": best.vit_cells / (best.ms_vit * 1e-3) / 1e9,
                      "thresholds": [0.3, 1e-4, 1e-5], "uniques": int(best.n_seq), "profiles": int(best.n_prof),
                      "note": "ms includes the two compactions around the per-profile launches; ~14 integer instructions "
                              "per cell (VIADDMNMX add-with-floor + max), rows in registers like fb_kernel"}))


if __name__ == "__main__":
    main()
