"""The CPU oracle against vectors produced by the unmodified reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import hudiff_oracle as O
from util import make_model, weight_checksum


@pytest.fixture(scope="module", params=["ab", "nb"])
def case(request, golden_dir):
    kind = request.param
    g = np.load(os.path.join(golden_dir, f"{kind}_release_b2.npz"))
    model, cfg = make_model(kind, seed=int(g["weight_seed"]))
    sd = model.state_dict()
    np.testing.assert_allclose(weight_checksum(sd), g["weight_checksum"], rtol=1e-12)
    return kind, cfg, sd, g


def _inputs(kind, g):
    tokens = torch.from_numpy(g["tokens"])
    region = torch.from_numpy(g["region"])
    chain = torch.from_numpy(g["chain"]) if kind == "ab" else None
    return tokens, region, chain


def test_logits_dropout_off(case):
    kind, cfg, sd, g = case
    tokens, region, chain = _inputs(kind, g)
    y = O.forward(kind, sd, dict(cfg, dropout=0.0), tokens, region, chain)
    ref = torch.from_numpy(g["logits_dropout_off"])
    assert y.shape == ref.shape
    assert float((y - ref).abs().max()) < 2e-5          # fp32 re-association only (dropout off)


def test_logits_dropout_live_replays_torch_stream(case):
    kind, cfg, sd, g = case
    tokens, region, chain = _inputs(kind, g)
    state = torch.get_rng_state()
    torch.manual_seed(int(g["dropout_seed"]))
    y = O.forward(kind, sd, cfg, tokens, region, chain, O.TorchStream())
    torch.set_rng_state(state)
    ref = torch.from_numpy(g["logits_dropout_live"])
    assert float((y - ref).abs().max()) < 1e-4            # fp32 re-association, amplified by the x2 dropout scale
    # the forward really is stochastic in eval mode
    assert float((ref - torch.from_numpy(g["logits_dropout_off"])).abs().max()) > 1e-2


def test_sampling_loop_tokens_exact(case):
    kind, cfg, sd, g = case
    tokens = torch.from_numpy(g["loop_tokens_in"])
    region = torch.from_numpy(g["region"])[:1].expand(tokens.shape[0], -1).clone()
    chain = torch.from_numpy(g["loop_chain"]) if kind == "ab" else None
    state = torch.get_rng_state()
    torch.manual_seed(int(g["loop_seed"]))
    out, sel = O.sample_loop(kind, sd, cfg, tokens, region, chain, torch.from_numpy(g["loop_loc"]),
                             O.TorchStream(), return_logits=True)
    torch.set_rng_state(state)
    assert torch.equal(out, torch.from_numpy(g["loop_tokens_out"]))
    assert float((sel - torch.from_numpy(g["loop_sel_logits"])).abs().max()) < 1e-4
    assert int((out == 22).sum()) == int((tokens == 22).sum()) - tokens.shape[0] * len(g["loop_loc"])


def test_recorded_stream_equals_torch_stream():
    """Pre-generating the masks/draws (what the B200 replay mode does) changes nothing."""
    kind = "nb"
    from hudiff_b200 import NB_RELEASE
    cfg = dict(NB_RELEASE, n_encoder_layers=1, dual_layers=1, cs_layers=1)
    model, cfg = make_model(kind, cfg, seed=2)
    sd = model.state_dict()
    from hudiff_b200.synthetic import synthetic
    tokens, region, _, order = synthetic(kind, 2, 3, seed=1)
    state = torch.get_rng_state()
    torch.manual_seed(9)
    a = O.sample_loop(kind, sd, cfg, tokens, region, None, order, O.TorchStream())
    torch.manual_seed(9)
    masks, exps = [], []
    for _ in range(3):
        masks += [torch.empty(2, 152, 256).bernoulli_(0.5)]
        masks += [torch.empty(2, 152, 512).bernoulli_(0.5)]
        exps += [torch.empty(2, 22).exponential_(1.0)]
    torch.set_rng_state(state)
    b = O.sample_loop(kind, sd, cfg, tokens, region, None, order, O.Recorded(masks, exps))
    assert torch.equal(a, b)
