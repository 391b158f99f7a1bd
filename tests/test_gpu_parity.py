"""Parity of the sm_100a path (through the C-ABI) with the CPU oracle and the golden vectors
made by the unmodified reference.  Run on the B200 box:  pytest tests -m gpu

Tolerances (written here on purpose, BASELINE.json north_star):
  * logits: max|d| / max|logit| <= 1e-3 for the default fp16-operand policy (TF32-grade, 11-bit
    significands, fp32 accumulate); <= 5e-5 for the split policy `fp16x3` (~fp32).
  * sampled tokens: exactly equal under the replayed torch CPU stream (fp16x3 policy).
"""
import os

import numpy as np
import pytest
import torch

from util import make_model, rel_err, weight_checksum

pytestmark = pytest.mark.gpu

from hudiff_b200 import AB_RELEASE, NB_RELEASE                       # noqa: E402
from hudiff_b200.synthetic import synthetic                           # noqa: E402
from oracle import hudiff_oracle as O                                 # noqa: E402

TOL_FP16 = 1e-3
TOL_X3 = 5e-5
SMALL = dict(n_encoder_layers=2, dual_layers=2, cs_layers=1)


def dev(t):
    return None if t is None else t.cuda()


@pytest.fixture(scope="module", params=["ab", "nb"])
def release(request, golden_dir):
    kind = request.param
    g = np.load(os.path.join(golden_dir, f"{kind}_release_b2.npz"))
    model, cfg = make_model(kind, seed=int(g["weight_seed"]))
    np.testing.assert_allclose(weight_checksum(model.state_dict()), g["weight_checksum"], rtol=1e-12)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    return kind, cfg, sd, model.cuda(), g


def _golden_inputs(kind, g):
    tokens = torch.from_numpy(g["tokens"])
    region = torch.from_numpy(g["region"])
    chain = torch.from_numpy(g["chain"]) if kind == "ab" else None
    return tokens, region, chain


@pytest.mark.parametrize("precision,tol", [("fp16", TOL_FP16), ("fp16x3", TOL_X3)])
def test_logits_vs_reference_golden(release, precision, tol):
    kind, cfg, sd, model, g = release
    tokens, region, chain = _golden_inputs(kind, g)
    model.set_precision(precision)
    y = model(dev(tokens), dev(region), dev(chain)).cpu()            # same call as the reference forward
    ref = torch.from_numpy(g["logits_dropout_off"])
    assert y.shape == ref.shape and y.dtype == torch.float32
    assert rel_err(y, ref) <= tol


def test_live_dropout_replay_vs_reference_golden(release):
    """The reference's eval-mode dropout, reproduced from the replayed torch CPU stream."""
    kind, cfg, sd, model, g = release
    tokens, region, chain = _golden_inputs(kind, g)
    model.set_precision("fp16x3")
    state = torch.get_rng_state()
    torch.manual_seed(int(g["dropout_seed"]))
    y = model(dev(tokens), dev(region), dev(chain), dropout="torch_cpu_replay").cpu()
    torch.set_rng_state(state)
    assert rel_err(y, torch.from_numpy(g["logits_dropout_live"])) <= TOL_X3


def test_sampling_loop_tokens_equal_reference_golden(release):
    """Literal reference loop (sample.py:499-513) vs model.sample: token-for-token."""
    kind, cfg, sd, model, g = release
    tokens = torch.from_numpy(g["loop_tokens_in"])
    B = tokens.shape[0]
    region = torch.from_numpy(g["region"])[:1].expand(B, -1).clone()
    chain = torch.from_numpy(g["loop_chain"]) if kind == "ab" else None
    loc = torch.from_numpy(g["loop_loc"])                              # one shared order, like the reference
    model.set_precision("fp16x3")
    state = torch.get_rng_state()
    torch.manual_seed(int(g["loop_seed"]))
    out, sel = model.sample(dev(tokens), dev(region), dev(chain), loc, rng="torch_cpu_replay",
                            dropout="reference", return_logits=True)
    torch.set_rng_state(state)
    assert out.dtype == tokens.dtype
    assert torch.equal(out.cpu(), torch.from_numpy(g["loop_tokens_out"]))
    assert rel_err(sel.cpu(), torch.from_numpy(g["loop_sel_logits"])) <= TOL_X3


@pytest.mark.parametrize("kind", ["ab", "nb"])
def test_per_sequence_order_vs_oracle(kind):
    """Independent sequences with their own sampling orders (the generalisation of the loop)."""
    base = AB_RELEASE if kind == "ab" else NB_RELEASE
    model, cfg = make_model(kind, dict(base, **SMALL), seed=4)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    tokens, region, chain, order = synthetic(kind, 3, 5, seed=9)
    state = torch.get_rng_state()
    torch.manual_seed(21)
    want, want_sel = O.sample_loop(kind, sd, cfg, tokens, region, chain, order, O.TorchStream(), return_logits=True)
    torch.manual_seed(21)
    model = model.cuda().set_precision("fp16x3")
    got, sel = model.sample(dev(tokens), dev(region), dev(chain), order, rng="torch_cpu_replay", return_logits=True)
    torch.set_rng_state(state)
    assert torch.equal(got.cpu(), want)
    assert rel_err(sel.cpu(), want_sel) <= TOL_X3
    # fp16 policy: per-step logits inside the 1e-3 tolerance on the same replayed stream
    torch.manual_seed(21)
    model.set_precision("fp16")
    _, sel16 = model.sample(dev(tokens), dev(region), dev(chain), order, rng="torch_cpu_replay", return_logits=True)
    torch.set_rng_state(state)
    assert rel_err(sel16[0].cpu(), want_sel[0]) <= TOL_FP16           # step 0: identical inputs on both sides


@pytest.mark.parametrize("kind,B", [("ab", 5), ("nb", 7)])
def test_tcgen05_gemm_equals_simt_reference(kind, B):
    """tcgen05/TMA kernels vs the CUDA-core reference GEMMs on ragged tile boundaries
    (B*152, B*144 and B*L are not multiples of the 128-row tile)."""
    base = AB_RELEASE if kind == "ab" else NB_RELEASE
    model, cfg = make_model(kind, dict(base, **SMALL), seed=5)
    model = model.cuda()
    tokens, region, chain, _ = synthetic(kind, B, 4, seed=2)
    for precision in ("fp16", "fp16x3"):
        model.set_precision(precision)
        eng = model.engine("cuda")
        eng.set_gemm_impl("tc")
        y_tc = model(dev(tokens), dev(region), dev(chain)).cpu()
        eng.set_gemm_impl("simt")
        y_ref = model(dev(tokens), dev(region), dev(chain)).cpu()
        eng.set_gemm_impl("tc")
        assert torch.isfinite(y_tc).all()
        # same operands; summation order differs, and in the single-fp16 policy a 1-ulp fp32 difference
        # can flip an fp16 operand rounding downstream (2^-11 relative on that element)
        assert rel_err(y_tc, y_ref) <= (5e-4 if precision == "fp16" else 2e-5), precision


def test_intermediate_streams_vs_oracle():
    """Localises an error to a stage: aa_encoder output and final residual stream."""
    model, cfg = make_model("ab", dict(AB_RELEASE, **SMALL), seed=6)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    tokens, region, chain, _ = synthetic("ab", 2, 4, seed=3)
    trace = {}
    O.forward("ab", sd, dict(cfg, dropout=0.0), tokens, region, chain, None, trace)
    model = model.cuda().set_precision("fp16x3")
    model(dev(tokens), dev(region), dev(chain))
    eng = model.engine("cuda")
    for name, key in (("E32", "E"), ("X32", "X")):
        buf = eng.debug_buffer(2, name).cpu().view(2, cfg["max_len"], -1)
        assert rel_err(buf, trace[key]) <= TOL_X3, name


def test_philox_mode_properties_full_size():
    """BASELINE config 2 shape (B=256, L=291, T=50 is bench territory; T=8 here): size-independent
    properties -- only the ordered slots change, every sample is one of the 22 classes, the run is
    deterministic in the seed, and sequences are independent of batch composition."""
    model, cfg = make_model("ab", seed=0)
    model = model.cuda().set_precision("fp16")
    B, T = 256, 8
    tokens, region, chain, order = synthetic("ab", B, T, seed=0)
    a = model.sample(dev(tokens), dev(region), dev(chain), order, rng="philox", seed=1).cpu()
    b = model.sample(dev(tokens), dev(region), dev(chain), order, rng="philox", seed=1).cpu()
    c = model.sample(dev(tokens), dev(region), dev(chain), order, rng="philox", seed=2).cpu()
    assert torch.equal(a, b) and not torch.equal(a, c)
    changed = a != tokens
    touched = torch.zeros_like(changed)
    touched.scatter_(1, order.long(), True)
    assert bool((changed <= touched).all())
    picked = torch.gather(a, 1, order.long())
    assert int(picked.min()) >= 0 and int(picked.max()) < 22
    assert int((a == 22).sum()) == int((tokens == 22).sum()) - B * T
    # independence of batch composition: with dropout off and replayed Exp(1) draws, the first 8
    # sequences sampled alone equal the same sequences sampled inside the batch of 256
    state = torch.get_rng_state()
    torch.manual_seed(3)
    full = model.sample(dev(tokens), dev(region), dev(chain), order, rng="torch_cpu_replay", dropout="off").cpu()
    torch.manual_seed(3)
    draws = torch.stack([torch.empty((B, 22)).exponential_(1.0) for _ in range(T)])
    torch.set_rng_state(state)
    sub = 8
    eng = model.engine("cuda")
    import ctypes as C
    from hudiff_b200 import capi
    tok = tokens[:sub].to(torch.int32).cuda().contiguous()
    reg = region[:sub].to(torch.int32).cuda().contiguous()
    chn = torch.cat([chain[:sub], chain[B:B + sub]]).to(torch.int32).cuda().contiguous()
    od = order[:sub].to(torch.int32).cuda().contiguous()
    ex = draws[:, :sub].contiguous().cuda()
    r = capi.HudiffRng(mode=1, dropout=0, seed=0, dropout_bits=None, exp_draws=ex.data_ptr())
    capi.check(eng.lib.hudiff_sample(eng.h, C.c_void_p(tok.data_ptr()), C.c_void_p(reg.data_ptr()),
                                     C.c_void_p(chn.data_ptr()), C.c_void_p(od.data_ptr()), sub, T, C.byref(r),
                                     C.c_void_p(None), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "hudiff_sample")
    torch.cuda.synchronize()
    assert torch.equal(tok.cpu().long(), full[:sub])


def test_host_buffer_entry_point_matches_device_entry_point():
    model, cfg = make_model("nb", dict(NB_RELEASE, **SMALL), seed=8)
    model = model.cuda().set_precision("fp16")
    tokens, region, _, order = synthetic("nb", 4, 6, seed=5)
    a = model.sample(dev(tokens), dev(region), None, order, rng="philox", seed=11).cpu()
    eng = model.engine("cuda")
    b = eng.sample_host(tokens.numpy(), region.numpy(), None, order.numpy(), seed=11)
    assert np.array_equal(a.numpy().astype(np.int32), b)


def test_edge_cases_and_errors():
    model, cfg = make_model("nb", dict(NB_RELEASE, **SMALL), seed=9)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    tokens, region, _, order = synthetic("nb", 1, 3, seed=6)
    with pytest.raises(RuntimeError):                 # no CPU path
        model(tokens, region, None)
    model = model.cuda().set_precision("fp16x3")
    # B = 1, T = 1, and a slot that is re-sampled twice (the reference's retry quirk re-samples slots)
    rep = torch.tensor([[int(order[0, 0]), int(order[0, 0]), int(order[0, 1])]])
    state = torch.get_rng_state()
    torch.manual_seed(2)
    want = O.sample_loop("nb", sd, cfg, tokens, region, None, rep, O.TorchStream())
    torch.manual_seed(2)
    got = model.sample(dev(tokens), dev(region), None, rep, rng="torch_cpu_replay").cpu()
    torch.set_rng_state(state)
    assert torch.equal(got, want)
    with pytest.raises(AssertionError):               # RoPE length assert of the reference (cross_attention.py:30)
        model(dev(tokens[:, :100]), dev(region[:, :100]), None)
    with pytest.raises(IndexError):
        model.sample(dev(tokens), dev(region), None, torch.tensor([[152]]))
    from hudiff_b200 import capi
    from hudiff_b200.engine import Engine
    bad = {k: v for k, v in model.state_dict().items() if k != "decoder.bias"}
    with pytest.raises(capi.HudiffError, match="decoder.bias"):
        Engine("nb", cfg, bad, torch.device("cuda", 0))


@pytest.mark.parametrize("env", [{"HUDIFF_LNFOLD": "1"}, {"HUDIFF_LNFOLD": "0"}, {"HUDIFF_2CTA": "0"}, {"HUDIFF_2CTA": "0", "HUDIFF_NSPLIT": "1"},
                                 {"HUDIFF_TAIL": "0"}, {"HUDIFF_ATTN": "mma"}, {"HUDIFF_ATTN": "tc2"}, {"HUDIFF_ATTN": "tc4"}, {"HUDIFF_ATTN_TAU": "-100"},
                                 {"HUDIFF_ATTN_TAU": "0", "HUDIFF_ATTN_SPLIT": "2"}, {"HUDIFF_ATTN_POLY": "0"}, {"HUDIFF_LNTAIL": "1", "HUDIFF_LNFOLD": "0"}, {"HUDIFF_PDL": "1"}, {"HUDIFF_L2HINT": "1", "HUDIFF_L2FLIP": "0", "HUDIFF_FUSE0": "0"}])
def test_optional_kernel_paths_keep_parity(env, golden_dir):
    """The environment-selected variants (folded LayerNorm, multicast instead of 2-CTA MMA, column-split
    LayerNorm tiles, no last-block shortcut, first-generation attention) meet the same tolerances: logits
    vs the reference golden vectors and the replayed sampling loop vs the oracle, token for token."""
    kind = "ab"
    g = np.load(os.path.join(golden_dir, f"{kind}_release_b2.npz"))
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)                              # read by hudiff_create
    try:
        model, cfg = make_model(kind, seed=int(g["weight_seed"]))
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        model = model.cuda()
        tokens, region, chain = _golden_inputs(kind, g)
        ref = torch.from_numpy(g["logits_dropout_off"])
        for precision, tol in (("fp16", TOL_FP16), ("fp16x3", TOL_X3)):
            model.set_precision(precision)
            y = model(dev(tokens), dev(region), dev(chain)).cpu()
            assert rel_err(y, ref) <= tol, (env, precision)
        # sampling loop (exercises the last-block shortcut unless disabled) against the oracle
        small = dict(AB_RELEASE, **SMALL)
        m2, cfg2 = make_model(kind, small, seed=4)
        sd2 = {k: v.clone() for k, v in m2.state_dict().items()}
        tk, rg, ch, order = synthetic(kind, 3, 5, seed=9)
        state = torch.get_rng_state()
        torch.manual_seed(21)
        want = O.sample_loop(kind, sd2, cfg2, tk, rg, ch, order, O.TorchStream())
        torch.manual_seed(21)
        got = m2.cuda().set_precision("fp16x3").sample(dev(tk), dev(rg), dev(ch), order, rng="torch_cpu_replay").cpu()
        torch.set_rng_state(state)
        assert torch.equal(got, want), env
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_last_block_shortcut_equals_full_path():
    """model.sample with the sampled-slot shortcut vs the same loop through full forwards (HUDIFF_TAIL=0):
    identical tokens and selected logits within the fp16x3 tolerance at a release-size model."""
    model_a, cfg = make_model("nb", seed=3)
    tokens, region, _, order = synthetic("nb", 6, 4, seed=12)
    outs = []
    for tail in ("1", "0"):
        old = os.environ.get("HUDIFF_TAIL")
        os.environ["HUDIFF_TAIL"] = tail
        try:
            m, _ = make_model("nb", seed=3)
            m = m.cuda().set_precision("fp16x3")
            state = torch.get_rng_state()
            torch.manual_seed(8)
            outs.append(m.sample(dev(tokens), dev(region), None, order, rng="torch_cpu_replay", dropout="off", return_logits=True))
            torch.set_rng_state(state)
        finally:
            if old is None:
                os.environ.pop("HUDIFF_TAIL", None)
            else:
                os.environ["HUDIFF_TAIL"] = old
    (ta, la), (tb, lb) = outs
    assert torch.equal(ta.cpu(), tb.cpu())
    assert rel_err(la.cpu(), lb.cpu()) <= TOL_X3
