"""The attention core alone (reference: model/encoder/cross_attention.py:131-173 after the q|k|v projection and RoPE)
through the C-ABI hook hudiff_debug_attention, for ragged sequence lengths and every kernel generation.

Oracle here is the plain fp32 softmax(q k^T) v of the fp16-rounded inputs (torch, on the GPU only to keep the test
fast); the tolerance is the fp16 rounding of the probabilities and of the output (2^-11 relative to the row scale)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

IMPLS = {"mma": 1, "tc2": 2, "tc3": 3, "tc4": 4}


def _run(qkv, B, L, A, H, impl, iters=0):
    from hudiff_b200 import capi
    lib = capi.load()
    out = torch.full((B * L, A), float("nan"), dtype=torch.float16, device=qkv.device)
    ms = C.c_float(0.0)
    rc = lib.hudiff_debug_attention(qkv.data_ptr(), out.data_ptr(), B, L, A, H, impl, iters, C.byref(ms))
    if rc != 0:
        raise capi.HudiffError(lib.hudiff_last_error().decode())
    return out, ms.value


def _reference(qkv, B, L, A, H):
    x = qkv.float().view(B, L, 3, H, 64)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)   # [B, H, L, 64]
    p = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    return (p @ v).transpose(1, 2).reshape(B * L, A)


def _inputs(B, L, H, seed, spread=1.0, late_peak=False):
    g = torch.Generator().manual_seed(seed)
    A = 64 * H
    x = torch.randn(B * L, 3 * A, generator=g) * 0.5
    x[:, :A] *= spread / 8.0 ** 0.5                 # q carries 1/sqrt(head dim)
    if late_peak:
        # rows whose largest scores sit at the END of the key axis (exercises the running-maximum refresh of the
        # online softmax in attention_tc3 and the two-half merge of attention_tc4 with very different maxima)
        xv = x.view(B, L, 3, H, 64)
        xv[:, -5:, 1] *= 6.0
    return x.half().cuda()


@pytest.mark.parametrize("impl", ["tc4", "tc3", "tc2"])
@pytest.mark.parametrize("L", [17, 31, 64, 100, 128, 129, 152, 160, 161, 255, 257, 291, 320])
def test_attention_core_ragged_lengths(impl, L):
    if impl == "tc2" and L not in (100, 152, 291, 320):
        pytest.skip("second-generation kernel: release lengths only")
    B, H = 5, 8
    qkv = _inputs(B, L, H, seed=L)
    ref = _reference(qkv, B, L, 64 * H, H)
    got, _ = _run(qkv, B, L, 64 * H, H, IMPLS[impl])
    assert torch.isfinite(got.float()).all(), "unwritten or non-finite output rows"
    err = (got.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 2e-3, (impl, L, err)


@pytest.mark.parametrize("impl", ["tc4", "tc3"])
@pytest.mark.parametrize("L", [152, 291])
def test_attention_core_late_maximum_and_wide_scores(impl, L):
    B, H = 3, 8
    qkv = _inputs(B, L, H, seed=7, spread=6.0, late_peak=True)
    ref = _reference(qkv, B, L, 64 * H, H)
    got, _ = _run(qkv, B, L, 64 * H, H, IMPLS[impl])
    err = (got.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 3e-3, (impl, L, err)


def test_attention_core_matches_validation_kernel_at_release_shape():
    """Every tcgen05 generation against the mma.sync validation kernel on the HuDiff-Ab shape (many items per SM, so the
    K / V double buffer, the Q ring and the tail-tile rotation all wrap around several times)."""
    B, L, H = 40, 291, 8
    qkv = _inputs(B, L, H, seed=3)
    base, _ = _run(qkv, B, L, 64 * H, H, IMPLS["mma"])
    for impl in ("tc2", "tc3", "tc4"):
        got, _ = _run(qkv, B, L, 64 * H, H, IMPLS[impl])
        err = (got.float() - base.float()).abs().max().item() / base.float().abs().max().item()
        assert err <= 2e-3, (impl, err)
