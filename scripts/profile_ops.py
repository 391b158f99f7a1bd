#!/usr/bin/env python
"""Per-launch device times of one denoiser forward (CUDA events around every launch, eager).

    python scripts/profile_ops.py [--kind ab|nb] [--batch B] [--dbg 0,1,2,4] [--out gpurun_out/ops.tsv]

`--dbg` repeats the forward with the GEMM kernel's timing experiments (HUDIFF_DBG bit mask:
1 = epilogue skipped, 2 = MMA issue skipped, 4 = TMA loads skipped); the results of those runs
are garbage by construction, only their durations matter.  Prints one row per distinct op label
with the mean time per launch, the launch count, TFLOP/s and the per-forward total.
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="ab")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--dbg", default="0")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ops.tsv"))
    ap.add_argument("--precision", default="fp16")
    args = ap.parse_args()
    from hudiff_b200 import AB_RELEASE, NB_RELEASE, AntiTFNetB200, NanoAntiTFNetB200
    from hudiff_b200.synthetic import synthetic
    kind = args.kind
    B = args.batch or (256 if kind == "ab" else 512)
    cfg = dict(AB_RELEASE if kind == "ab" else NB_RELEASE)
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    model = (AntiTFNetB200 if kind == "ab" else NanoAntiTFNetB200)(**cfg).to(dev).set_precision(args.precision)
    eng = model.engine(dev)
    tokens, region, chain, order = synthetic(kind, B, 4, seed=1)
    d_tok, d_reg = tokens.to(dev), region.to(dev)
    d_chn = chain.to(dev) if chain is not None else None
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    if os.path.exists(args.out):
        os.remove(args.out)
    import ctypes as C
    gclk = (C.c_ulonglong * 32)()
    for dbg in args.dbg.split(","):
        os.environ["HUDIFF_DBG"] = dbg
        os.environ.pop("HUDIFF_PROFILE_DUMP", None)
        eng.profile_step(d_tok, d_reg, d_chn, dropout="reference", seed=3)       # warm
        os.environ["HUDIFF_PROFILE_DUMP"] = args.out
        eng.lib.hudiff_debug_gemm_clocks(gclk, 1)
        for _ in range(3):
            eng.profile_step(d_tok, d_reg, d_chn, dropout="reference", seed=3)
        if int(dbg) & 8:
            eng.lib.hudiff_debug_gemm_clocks(gclk, 1)
            for e in range(4):
                n = gclk[e * 8]
                if n:
                    print(f"gemm epilogue clocks dbg={dbg} epi{e}: tiles {n}",
                          {k: round(gclk[e * 8 + 1 + i] / n) for i, k in enumerate(["wait_acc", "ln_stats", "tmem_to_smem", "phase2"])},
                          "cycles/tile (warp 2 of CTA 0)")
    os.environ["HUDIFF_DBG"] = "0"
    agg = collections.OrderedDict()
    for line in open(args.out):
        dbg, idx, label, ms, gf = line.rstrip("\n").split("\t")
        a = agg.setdefault((dbg, label), [0.0, 0, 0.0])
        a[0] += float(ms); a[1] += 1; a[2] += float(gf)
    print(f"{'dbg':>3} {'op':<34} {'n/fwd':>5} {'us/launch':>10} {'TFLOP/s':>8} {'ms/fwd':>8}")
    totals = collections.defaultdict(float)
    for (dbg, label), (ms, n, gf) in agg.items():
        per = ms / n
        tf = (gf / n) / per if per > 0 and gf > 0 else 0.0
        print(f"{dbg:>3} {label:<34} {n // 3:>5} {per * 1e3:>10.1f} {tf:>8.1f} {ms / 3:>8.3f}")
        totals[dbg] += ms / 3
    for dbg, t in totals.items():
        print(f"dbg={dbg}: forward total {t:.3f} ms")
    import ctypes as C
    clk = (C.c_longlong * 16)()
    rc_clk = eng.lib.hudiff_debug_attn_clocks(clk)
    if rc_clk == 0 and clk[0] > 0:
        n = clk[0]
        names = ["wait_S", "tmem_ld", "max_xchg", "exp_P", "readback", "fence_arrive"]
        print("attention_tc2 CTA0 cycles/tile:", {k: round(clk[i + 1] / n) for i, k in enumerate(names)}, "tiles", n)
    if clk[8] > 0:
        n = clk[8]
        print("attention_tc3 CTA0 slot 0 cycles/tile:", {"total": round(clk[9] / n), "wait_S": round(clk[10] / n), "wait_O": round(clk[11] / n)}, "tiles", n)


if __name__ == "__main__":
    main()
