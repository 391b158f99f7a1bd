"""Bring-up script for the GPU box (not a pytest): runs each check in its own process so that a
trapped kernel cannot poison the next one.   python tests/gpu_bringup.py [stage ...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def stage(name):
    import torch
    from oracle import hudiff_oracle as O
    from hudiff_b200 import AB_RELEASE, NB_RELEASE
    from hudiff_b200.synthetic import synthetic
    from util import make_model, rel_err

    kind, cfgname, impl, prec, what = name.split(":")
    base = AB_RELEASE if kind == "ab" else NB_RELEASE
    cfg = dict(base)
    if cfgname == "small":
        cfg.update(n_encoder_layers=2, dual_layers=2, cs_layers=1)
    elif cfgname == "tiny":
        cfg.update(n_encoder_layers=1, dual_layers=1, cs_layers=1)
    B = 2
    model, cfg = make_model(kind, cfg, seed=0)
    sd = model.state_dict()
    tokens, region, chain, order = synthetic(kind, B, 6, seed=3)
    model = model.to("cuda").set_precision(prec)
    eng = model.engine("cuda")
    eng.set_gemm_impl(impl)
    dev = lambda t: None if t is None else t.cuda()
    out = {"stage": name}
    if what == "fwd":
        trace = {}
        ref = O.forward(kind, sd, dict(cfg, dropout=0.0), tokens, region, chain, None, trace)
        t0 = time.time()
        y = eng.forward(dev(tokens), dev(region), dev(chain), dropout="off").cpu()
        out["secs"] = round(time.time() - t0, 3)
        out["logits_rel_err"] = rel_err(y, ref)
        out["finite"] = bool(torch.isfinite(y).all())
        L = cfg["max_len"]
        for nm, key in (("E32", "E"), ("X32", "X")):
            buf = eng.debug_buffer(B, nm).cpu().view(B, L, -1)
            out[nm + "_rel_err"] = rel_err(buf, trace[key])
    elif what == "drop":
        torch.manual_seed(11)
        ref = O.forward(kind, sd, cfg, tokens, region, chain, O.TorchStream())
        torch.manual_seed(11)
        y = eng.forward(dev(tokens), dev(region), dev(chain), dropout="torch_cpu_replay").cpu()
        out["logits_rel_err"] = rel_err(y, ref)
    elif what == "loop":
        torch.manual_seed(7)
        ref_tok, ref_sel = O.sample_loop(kind, sd, cfg, tokens, region, chain, order, O.TorchStream(), return_logits=True)
        torch.manual_seed(7)
        tok, sel = model.sample(dev(tokens), dev(region), dev(chain), order, rng="torch_cpu_replay",
                                dropout="reference", return_logits=True)
        out["token_mismatch"] = int((tok.cpu() != ref_tok).sum())
        out["sel_logits_rel_err"] = rel_err(sel.cpu(), ref_sel)
        out["launches"] = eng.last_launch_count()
    print("RESULT " + json.dumps(out), flush=True)


DEFAULT = [
    "nb:tiny:simt:fp16x3:fwd", "nb:tiny:tc:fp16x3:fwd", "nb:tiny:tc:fp16:fwd",
    "ab:tiny:simt:fp16x3:fwd", "ab:tiny:tc:fp16x3:fwd", "ab:tiny:tc:fp16:fwd",
    "ab:small:tc:fp16x3:drop", "ab:small:tc:fp16x3:loop", "nb:small:tc:fp16x3:loop",
    "ab:full:tc:fp16:fwd", "ab:full:tc:fp16x3:fwd",
]

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--stage":
        stage(sys.argv[2])
        sys.exit(0)
    stages = sys.argv[1:] or DEFAULT
    for st in stages:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", st], capture_output=True,
                               text=True, timeout=300)
            lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                print(lines[-1], "| wall", round(time.time() - t0, 1), flush=True)
            else:
                print(f"FAILED {st} rc={p.returncode}\n--- stdout\n{p.stdout[-1500:]}\n--- stderr\n{p.stderr[-3000:]}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"TIMEOUT {st}", flush=True)
