#!/usr/bin/env python
"""Benchmark of the HuDiff sampling hot path (BASELINE.json: humanized sequences / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE complete T-step denoising loop over one batch of synthetic sequences per GPU
(HuDiff-Ab, B=256 independent paired sequences on the 152+139 slot grid, T=50 masked slots each;
SURVEY.md §8d config 2).  `value` is whole-job sequences/s with inputs resident in HBM (device
pointers into hudiff_sample); `e2e` is the same loop through the host-buffer C-ABI call
(hudiff_sample_host: pinned host tokens in, tokens out).  Random-init weights of the released
architecture, the reference's live eval-mode dropout included (in-kernel Philox), fp16-operand
tensor-core policy (TF32-grade significands, fp32 accumulate; meets the 1e-3 logits tolerance).

`--impl reference` times the reference algorithm on the host CPU cores: the reference is pure
Python/PyTorch and cannot travel to the GPU box, so the oracle port (oracle/hudiff_oracle.py,
pinned against the reference by tests/golden) runs the same loop on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# reference-equivalent algorithmic work of one denoiser forward (FlopCounterMode)
GFLOP_PER_SEQ_STEP = {"ab": 18.4897, "nb": 5.7865}
T_DEFAULT = 50


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"burst": float(p["bf16_tflops"]), "sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "hbm": float(p["hbm_gbs"]), "source": "MEASURED_PEAKS.json"}
    except Exception:
        return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


NCU_GEMM_FILES = ("r02b_ncu_full_gemm.json", "r02_ncu_full_gemm.json", "r01c_ncu_full_gemm.json")


def ncu_traffic_per_gemm_launch():
    """(mean DRAM bytes (read + write) per gemm_tc_kernel launch, file) from the committed `ncu --set full` capture
    (profiles/r02b_ncu_full_gemm.json: the 67 GEMM launches of one sampling step), or (None, None)."""
    for name in NCU_GEMM_FILES:
        v = _ncu_traffic(os.path.join(ROOT, "profiles", name))
        if v is not None:
            return v, name
    return None, None


def _ncu_traffic(path):
    try:
        with open(path) as f:
            recs = json.load(f)
        tot = []
        for r in recs:
            b = 0.0
            for k, v in r.items():
                if k.startswith("dram__bytes_read.sum") or k.startswith("dram__bytes_write.sum"):
                    unit = k.split("[")[-1].rstrip("]").lower()
                    b += float(v) * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
            tot.append(b)
        return float(np.mean(tot)) if tot else None
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [v for v in sm if mx and v > 0.3 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_reference_rate(kind, cfg, T, batch, n_forwards, warm):
    """seq/s of the CPU oracle port: `n_forwards` timed loop steps (forward + resample) on `batch`
    sequences with every host thread, extrapolated to T steps (every step costs the same)."""
    from hudiff_b200 import AntiTFNetB200, NanoAntiTFNetB200
    from hudiff_b200.synthetic import synthetic
    from oracle import hudiff_oracle as O
    torch.manual_seed(0)
    model = (AntiTFNetB200 if kind == "ab" else NanoAntiTFNetB200)(**cfg)
    sd = model.state_dict()
    tokens, region, chain, order = synthetic(kind, batch, warm + n_forwards, seed=0)
    stream = O.TorchStream()
    rows = torch.arange(batch)
    times = []
    with torch.no_grad():
        for t in range(warm + n_forwards):
            t0 = time.perf_counter()
            logits = O.forward(kind, sd, cfg, tokens, region, chain, stream)
            pos = order[:, t].long()
            tokens[rows, pos] = O.sample_step(logits[rows, pos], stream.exponential((batch, 22)))
            if t >= warm:
                times.append(time.perf_counter() - t0)
    per_forward = float(np.mean(times))
    return batch / (T * per_forward), per_forward


def run_reference(args, rank, world):
    if rank != 0:
        return
    from hudiff_b200 import AB_RELEASE, NB_RELEASE
    kind = args.kind
    cfg = dict(AB_RELEASE if kind == "ab" else NB_RELEASE)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch, fw = args.cpu_batch, args.cpu_forwards
    for _ in range(args.warmup):
        cpu_reference_rate(kind, cfg, args.T, batch, 1, 0)
    vals, ms = [], []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        v, _ = cpu_reference_rate(kind, cfg, args.T, batch, fw, 0)
        vals.append(v)
        ms.append((time.perf_counter() - t0) * 1e3)
    value = float(np.mean(vals))
    sample = (f"{fw} loop steps (forward + resample) on {batch} sequences per bench step, all {cores} host threads, "
              f"extrapolated x{args.T}/{fw} (step-homogeneous loop)")
    line = {
        "impl": "reference", "metric": "humanized sequences/sec", "value": value, "unit": "seq/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(ms)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the job's config (the arm measures the ONE host's CPU on a bounded sample of that workload; it does not shard)
        "config": workload_config(args, kind, args.batch_per_gpu, max(world, args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "seq/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, kind, B, world):
    name = "HuDiff-Ab" if kind == "ab" else "HuDiff-Nb"
    L = 291 if kind == "ab" else 152
    mask = ""
    if kind == "nb" and args.inpaint:
        mask = (" from the inpaint mask (INPAINT_HEAVY_CDR_INDEX: 87 candidate slots, 85 after the two framework gaps of the "
                "synthetic VHH template, which the reference never samples; BASELINE config 5 names T=100, unreachable "
                "with the reference's masks)")
    return {"workload": f"{name} sampling loop, batch={B}/GPU synthetic L={L}"
                        f"{' (152 heavy + 139 light)' if kind == 'ab' else ''} sequences, T={args.T} masked slots{mask}, "
                        f"released architecture (random-init), live eval dropout",
            "batch_per_gpu": B, "global_batch": B * world, "seq_len": L, "T": args.T, "precision": args.precision,
            "rng": "philox", "parallelism": f"batch-shard x{world}",
            "l2": "per-step activation working set (>1 GB at B=256) exceeds the 126 MB L2; no explicit flush"}


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) to stderr."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)                                          # libraries that print to fd 1 now land on stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kind", default="ab", choices=["ab", "nb"])
    ap.add_argument("--batch-per-gpu", type=int, default=None)
    ap.add_argument("--T", type=int, default=T_DEFAULT)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp16x3"])
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--cpu-forwards", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inpaint", action="store_true", help="Nb: inpaint mask (nanosample.py --inpaint_sample True), T <= 87")
    ap.add_argument("--no-x3", action="store_true", help="skip the extra timing of the token-exact fp16x3 policy")
    args = ap.parse_args()
    if args.batch_per_gpu is None:
        args.batch_per_gpu = 256 if args.kind == "ab" else 512
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3                                    # timing rule: at least 3 warm-up steps

    from hudiff_b200 import dist as hd
    rank, local_rank, world = hd.env_world()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from hudiff_b200 import AB_RELEASE, NB_RELEASE, AntiTFNetB200, NanoAntiTFNetB200
    from hudiff_b200.synthetic import synthetic

    rank, local_rank, world = hd.init_from_env("nccl")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    kind, B, T = args.kind, args.batch_per_gpu, args.T
    cfg = dict(AB_RELEASE if kind == "ab" else NB_RELEASE)

    # weights: seeded random init of the released architecture; ONE NCCL broadcast at start
    torch.manual_seed(0)
    model = (AntiTFNetB200 if kind == "ab" else NanoAntiTFNetB200)(**cfg).to(device)
    t0 = time.perf_counter()
    bcast_bytes = hd.broadcast_model(model, src=0)
    torch.cuda.synchronize()
    startup_s = time.perf_counter() - t0
    model.set_precision(args.precision)
    eng = model.engine(device)

    from hudiff_b200.synthetic import synthetic_nb
    if kind == "nb" and args.inpaint:
        tokens, region, chain, order = synthetic_nb(B, T, seed=1000 + rank, inpaint=True)
    else:
        tokens, region, chain, order = synthetic(kind, B, T, seed=1000 + rank)
    d_tok, d_reg, d_ord = tokens.to(device), region.to(device), order.to(device)
    d_chn = chain.to(device) if chain is not None else None
    seed = 12345                          # one key for the job: rank r draws the streams of sequences [r*B, (r+1)*B)
    seq_offset = rank * B

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        out = model.sample(d_tok, d_reg, d_chn, d_ord, rng="philox", seed=seed, dropout="reference",
                           seq_offset=seq_offset, check=False)
        return hd.gather_tokens(out, dst=0)            # the final gather is part of the step

    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        result = one_step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item())
    eng.check_error()                     # device watchdogs / range checks of the timed calls (synchronises)
    launches = eng.last_launch_count() * args.steps
    value = world * B / (ms_per_step / 1e3)

    # ---- e2e: the same loop through the host-buffer C-ABI entry point (H2D + D2H inside) ----------
    h_tok = tokens.to(torch.int32).pin_memory().numpy()
    h_reg = region.to(torch.int32).pin_memory().numpy()
    h_chn = chain.to(torch.int32).pin_memory().numpy() if chain is not None else None
    h_ord = order.to(torch.int32).pin_memory().numpy()
    eng.sample_host(h_tok, h_reg, h_chn, h_ord, seed=seed, seq_offset=seq_offset)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_out = eng.sample_host(h_tok, h_reg, h_chn, h_ord, seed=seed, seq_offset=seq_offset)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / args.steps], device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B / float(e2e_s.item())
    h2d = (h_tok.nbytes + h_reg.nbytes + h_ord.nbytes + (h_chn.nbytes if h_chn is not None else 0))
    d2h = host_out.nbytes

    # ---- the token-exact operand policy (fp16x3: the mode the token-for-token parity tests run) timed beside it
    x3 = None
    if args.precision == "fp16" and not args.no_x3:
        model.set_precision("fp16x3")
        eng3 = model.engine(device)
        for _ in range(2):
            model.sample(d_tok, d_reg, d_chn, d_ord, rng="philox", seed=seed, dropout="reference", seq_offset=seq_offset, check=False)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n3 = max(2, min(args.steps, 3))
        e0.record()
        for _ in range(n3):
            model.sample(d_tok, d_reg, d_chn, d_ord, rng="philox", seed=seed, dropout="reference", seq_offset=seq_offset, check=False)
        e1.record()
        barrier()
        eng3.check_error()
        ms3 = torch.tensor([e0.elapsed_time(e1) / n3], device=device)
        if world > 1:
            dist.all_reduce(ms3, op=dist.ReduceOp.MAX)
        x3 = {"value": world * B / (float(ms3.item()) / 1e3), "unit": "seq/s", "ms_per_step": float(ms3.item()), "steps": n3,
              "note": "same workload with split-fp16 operands (3 MMA passes, ~fp32): the policy under which the sampling loop is "
                      "token-for-token equal to the reference (tests/test_gpu_parity.py, test_gpu_statistics.py)"}
        model.set_precision(args.precision)
        eng = model.engine(device)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (tcgen05 GEMM / dilated conv), events on its launch stream
    peaks = load_peaks()
    prof = eng.profile_step(d_tok, d_reg, d_chn, dropout="reference", seed=seed)
    prof = eng.profile_step(d_tok, d_reg, d_chn, dropout="reference", seed=seed)      # second run: warm
    g = prof["gemm_tc"]
    achieved = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    step_fwd_ms = sum(v["ms"] for v in prof.values())
    traffic, traffic_file = ncu_traffic_per_gemm_launch()
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peaks["sustained"], "unit": "TFLOP/s",
        "frac": achieved / peaks["sustained"], "traffic": traffic,
        "traffic_note": f"mean DRAM bytes (read+write) per GEMM launch, ncu --set full of the GEMM launches of one step (profiles/{traffic_file})",
        "timing": "per-kernel CUDA events recorded INSIDE one CUDA-graph launch of the forward (hudiff_profile_step): the "
                  "breakdown sums to ms_per_step / T",
        "kernel": "gemm_tc_kernel (all launches of one denoiser forward, aggregated)",
        "launches_per_forward": g["launches"], "avg_launch_ms": g["ms"] / max(1, g["launches"]),
        "share_of_forward": g["ms"] / step_fwd_ms if step_fwd_ms else None,
        "peak_source": f"{peaks['source']} bf16 sustained (kernel timed inside a long step); burst {peaks['burst']}",
        "frac_of_burst": achieved / peaks["burst"],
        "forward_breakdown_ms": {k: round(v["ms"], 4) for k, v in prof.items()},
        "whole_step_tflops_reference_equivalent": value * GFLOP_PER_SEQ_STEP[kind] * T / 1e3 / world,
    }

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        v, per_fwd = cpu_reference_rate(kind, cfg, T, args.cpu_batch, 3, 1)
        cpu_baseline = {"value": v, "unit": "seq/s", "cores": cores, "kind": "port",
                        "sample": f"oracle port (PyTorch-CPU fp32, live dropout), 3 timed loop steps on {args.cpu_batch} "
                                  f"sequences after 1 warm-up ({per_fwd:.2f} s/step), extrapolated to T={T}"}

    line = {
        "metric": "humanized sequences/sec", "value": value, "unit": "seq/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16 operands / f32 accumulate" if args.precision == "fp16" else "split-f16 x3 / f32 accumulate",
        "data": "synthetic", "config": workload_config(args, kind, B, world),
        "e2e": {"value": e2e_value, "unit": "seq/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "fp16x3": x3,
        "startup": {"weight_broadcast_bytes": int(bcast_bytes), "seconds": startup_s},
        "gathered_rows": int(result.shape[0]) if result is not None else None,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
