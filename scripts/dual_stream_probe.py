#!/usr/bin/env python
"""Feasibility probe: two half-batches on two engines / two CUDA streams, persistent grids capped to half
the SMs each (HUDIFF_MAX_SMS), vs one engine on the full batch.  Prints seq/s for both."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from hudiff_b200 import AB_RELEASE, AntiTFNetB200
    from hudiff_b200.synthetic import synthetic
    B, T = int(os.environ.get("PROBE_B", 256)), 50
    cap = os.environ.get("PROBE_CAP", "74")
    dev = torch.device("cuda", 0)
    tokens, region, chain, order = synthetic("ab", B, T, seed=1)

    def build():
        torch.manual_seed(0)
        return AntiTFNetB200(**AB_RELEASE).to(dev).set_precision("fp16")

    def run_single(m, n=3):
        d = [t.to(dev) for t in (tokens, region, chain, order)]
        for _ in range(2):
            m.sample(*d, rng="philox", seed=1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            m.sample(*d, rng="philox", seed=1)
        torch.cuda.synchronize()
        return B * n / (time.perf_counter() - t0)

    os.environ.pop("HUDIFF_MAX_SMS", None)
    print("single engine, all SMs:", round(run_single(build()), 1), "seq/s")

    os.environ["HUDIFF_MAX_SMS"] = cap
    h = B // 2
    halves = []
    for k in range(2):
        sl = slice(k * h, (k + 1) * h)
        ch = torch.cat([chain[:B][sl], chain[B:][sl]])
        halves.append([t.to(dev) for t in (tokens[sl], region[sl], ch, order[sl])])
    models = [build(), build()]
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]

    def step():
        for m, d, s in zip(models, halves, streams):
            with torch.cuda.stream(s):
                m.sample(*d, rng="philox", seed=1)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    print(f"two engines x {h} sequences, two streams, {cap} SMs each:", round(B * n / (time.perf_counter() - t0), 1), "seq/s")


if __name__ == "__main__":
    main()
