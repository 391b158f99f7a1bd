"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

    python oracle/make_golden.py

Runs the reference's `AntiTFNet` / `NanoAntiTFNet` (imported from /root/reference with the
stubs of `oracle/ref_stubs.py`) and the literal loop body of `antibody_scripts/sample.py:499-513`
/ `nanobody_scripts/nanosample.py:316-329` on small seeded inputs and stores inputs + outputs.
Weights are NOT stored: they are the default nn init under `torch.manual_seed(WEIGHT_SEED)`,
which `hudiff_b200.modules` reproduces bit-for-bit (checked here and by a checksum in the file).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hudiff_b200 import AB_RELEASE, NB_RELEASE, AntiTFNetB200, NanoAntiTFNetB200  # noqa: E402
from hudiff_b200.synthetic import synthetic  # noqa: E402
from oracle import ref_stubs  # noqa: E402

WEIGHT_SEED = 0
INPUT_SEED = 3
DROPOUT_SEED = 11
LOOP_SEED = 7
B = 2
T = 6


def weight_checksum(sd) -> np.ndarray:
    s = 0.0
    a = 0.0
    for k in sorted(sd):
        v = sd[k]
        v = torch.view_as_real(v) if v.is_complex() else v
        s += float(v.double().sum())
        a += float(v.double().abs().sum())
    return np.array([s, a])


def reference_loop(model, tokens, region, chain, loc):
    """Literal body of sample.py:501-513 (Ab) / nanosample.py:318-329 (Nb)."""
    tokens = tokens.clone()
    sel_logits = []
    with torch.no_grad():
        for i in loc:
            if chain is not None:
                pred = model(tokens, region, chain)
            else:
                pred = model(tokens, region, H_chn_type=None)
            h_l_pred = pred[:, i, :23 - 1]
            h_l_soft = torch.nn.functional.softmax(h_l_pred, dim=1)
            h_l_sample = torch.multinomial(h_l_soft, num_samples=1)
            tokens[:, i] = h_l_sample.squeeze()
            sel_logits.append(pred[:, i, :].clone())
    return tokens, torch.stack(sel_logits, 0)


def main():
    AntiTFNet, NanoAntiTFNet = ref_stubs.load_models()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for kind, cfg, Ref, Mine in (("ab", AB_RELEASE, AntiTFNet, AntiTFNetB200),
                                 ("nb", NB_RELEASE, NanoAntiTFNet, NanoAntiTFNetB200)):
        torch.manual_seed(WEIGHT_SEED)
        ref = Ref(**cfg).eval()
        torch.manual_seed(WEIGHT_SEED)
        mine = Mine(**cfg)
        sd = ref.state_dict()
        for k, v in mine.state_dict().items():
            assert torch.equal(v, sd[k]), f"mirror module diverged at {k}"
        ref_nodrop = Ref(**dict(cfg, dropout=0.0)).eval()
        ref_nodrop.load_state_dict(sd)

        tokens, region, chain, order = synthetic(kind, B, T, seed=INPUT_SEED)
        with torch.no_grad():
            args = (tokens, region, chain) if kind == "ab" else (tokens, region, None)
            logits_off = ref_nodrop(*args)
            torch.manual_seed(DROPOUT_SEED)
            logits_live = ref(*args)

        # reference loop semantics: B replicas of ONE sequence, one shared order (sample.py:172-176,498)
        rep_tokens = tokens[:1].expand(B, -1).clone()
        rep_region = region[:1].expand(B, -1).clone()
        rep_chain = None
        if kind == "ab":
            rep_chain = torch.stack([chain[0], chain[B]]).view(-1, 1).repeat(1, B).view(-1)
        loc = order[0].numpy().astype(np.int64)
        torch.manual_seed(LOOP_SEED)
        loop_tokens, loop_logits = reference_loop(ref, rep_tokens, rep_region, rep_chain, loc)

        np.savez_compressed(
            os.path.join(out_dir, f"{kind}_release_b{B}.npz"),
            weight_seed=WEIGHT_SEED, weight_checksum=weight_checksum(sd),
            tokens=tokens.numpy(), region=region.numpy(),
            chain=(chain.numpy() if chain is not None else np.zeros(0, dtype=np.int64)),
            order=order.numpy(),
            logits_dropout_off=logits_off.numpy(),
            dropout_seed=DROPOUT_SEED, logits_dropout_live=logits_live.numpy(),
            loop_seed=LOOP_SEED, loop_loc=loc, loop_tokens_in=rep_tokens.numpy(),
            loop_chain=(rep_chain.numpy() if rep_chain is not None else np.zeros(0, dtype=np.int64)),
            loop_tokens_out=loop_tokens.numpy(), loop_sel_logits=loop_logits.numpy(),
            torch_version=str(torch.__version__))
        print(kind, "golden written; logits max", float(logits_off.abs().max()),
              "changed tokens", int((loop_tokens != rep_tokens).sum()))


if __name__ == "__main__":
    main()
