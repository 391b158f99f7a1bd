#!/usr/bin/env python
"""Attention core alone (C-ABI hook hudiff_debug_attention): device time per launch of every kernel generation on the
BASELINE shapes, random fp16 q|k|v resident in HBM, mean of `--iters` back-to-back launches (CUDA events).

    python scripts/attention_bench.py [--iters 20] [--impl tc2,tc3,tc4]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

IMPLS = {"mma": 1, "tc2": 2, "tc3": 3, "tc4": 4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--impl", default="tc2,tc3,tc4")
    ap.add_argument("--shapes", default="ab:256:291,nb:512:152", help="kind:B:L,...")
    args = ap.parse_args()
    from hudiff_b200 import capi
    lib = capi.load()
    rows = []
    for kind, B, L in [(t.split(":")[0], int(t.split(":")[1]), int(t.split(":")[2])) for t in args.shapes.split(",")]:
        H, A = 8, 512
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B * L, 3 * A, generator=g) * 0.5
        x[:, :A] /= 8.0 ** 0.5
        qkv = x.half().cuda()
        out = torch.empty(B * L, A, dtype=torch.float16, device="cuda")
        for impl in args.impl.split(","):
            ms = C.c_float(0.0)
            rc = lib.hudiff_debug_attention(qkv.data_ptr(), out.data_ptr(), B, L, A, H, IMPLS[impl], args.iters, C.byref(ms))
            if rc != 0:
                raise SystemExit(lib.hudiff_last_error().decode())
            flops = 4.0 * B * H * L * L * 64
            rows.append({"kind": kind, "B": B, "L": L, "impl": impl, "us_per_launch": round(ms.value * 1e3, 1),
                         "tflops": round(flops / ms.value / 1e9, 1)})
            print(json.dumps(rows[-1]))


if __name__ == "__main__":
    main()
