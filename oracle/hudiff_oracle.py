"""CPU oracle for the HuDiff sampling hot path  --  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional PyTorch-CPU fp32, the arithmetic of the
reference's denoiser forward and of its order-agnostic sampling loop.  It is the
checker the CUDA path is compared with; it is never imported by the product
package `hudiff_b200` (only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may use it).

Parity pinning: the reference ships no tests or golden vectors, so
this oracle is pinned against the reference ITSELF, imported in the build
container with stubbed third-party packages (`oracle/ref_stubs.py`); the outputs
of that run are committed as `tests/golden/*.npz` by `oracle/make_golden.py` and
re-checked on every CPU test run (`tests/test_oracle_golden.py`).

Citations are relative to /root/reference.

Third-party arithmetic that is not in the reference tree: `sequence-models`
(unpinned, environment.yaml:21) provides ByteNetBlock / MaskedConv1d /
PositionFeedForward.  Published algorithm restated here:
  PositionFeedForward(d_in, d_out) = Conv1d(d_in, d_out, 1) on (N, L, C) input
  MaskedConv1d(c, c, k, dilation)  = Conv1d with zero 'same' padding d*(k-1)//2
  ByteNetBlock: x + seq2(conv(seq1(x))),
      seq1 = LN(d_in), act, PFF(d_in->d_h), LN(d_h), act
      seq2 = LN(d_h),  act, PFF(d_h->d_out)
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

H_LEN = 152   # model/encoder/model.py:185  (aa_h_length)
L_LEN = 139   # model/encoder/model.py:185  (aa_l_length)
N_SAMPLED_TOKENS = 22  # sample.py:510  logits[:, i, :len(toks)-1]


# --------------------------------------------------------------------------- #
# dropout sources: the reference calls F.dropout functionally (training=True)  #
# in eval mode (model/encoder/model.py:176-178,295-296,302-303;               #
# model/nanoencoder/model.py:168-169,264-265)                                  #
# --------------------------------------------------------------------------- #
class DropoutOff:
    """No dropout (config.model.dropout == 0, or deterministic logits tests)."""

    def keep(self, shape, p):
        return None

    def exponential(self, shape):
        raise RuntimeError("DropoutOff has no sampling stream")


class TorchStream:
    """Draws from torch's global CPU generator in the reference's call order.

    F.dropout(x, p) on CPU == x * empty_like(x).bernoulli_(1-p).div_(1-p)
    torch.multinomial(p, 1) == argmax(p / empty_like(p).exponential_(1))
    (aten/src/ATen/native/Dropout.cpp, Distributions.cpp).
    """

    def keep(self, shape, p):
        return torch.empty(shape, dtype=torch.float32).bernoulli_(1.0 - p)

    def exponential(self, shape):
        return torch.empty(shape, dtype=torch.float32).exponential_(1.0)


class Recorded:
    """Replays explicit keep-masks / Exp(1) draws (lists consumed in order)."""

    def __init__(self, masks: Sequence[Tensor], exps: Sequence[Tensor] = ()):
        self.masks = list(masks)
        self.exps = list(exps)
        self.mi = 0
        self.ei = 0

    def keep(self, shape, p):
        m = self.masks[self.mi]
        self.mi += 1
        assert tuple(m.shape) == tuple(shape), (m.shape, shape)
        return m.to(torch.float32)

    def exponential(self, shape):
        e = self.exps[self.ei]
        self.ei += 1
        assert tuple(e.shape) == tuple(shape)
        return e


def _apply_dropout(x: Tensor, p: float, src) -> Tensor:
    keep = src.keep(tuple(x.shape), p)
    if keep is None:
        return x
    # noise.div_(1-p) then input * noise  (ATen Dropout.cpp _dropout_impl)
    return x * keep.div(1.0 - p)


# --------------------------------------------------------------------------- #
# building blocks                                                              #
# --------------------------------------------------------------------------- #
def _act(name: str) -> Callable[[Tensor], Tensor]:
    # nn.GELU() is the exact erf form; ByteNetBlock picks ReLU unless 'gelu'.
    return F.gelu if name == "gelu" else F.relu


def _ln(x: Tensor, sd, key: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], 1e-5)


def _pff(x: Tensor, sd, key: str) -> Tensor:
    # PositionFeedForward == Conv1d(k=1) == per-row linear
    w = sd[key + ".conv.weight"]          # [d_out, d_in, 1]
    return F.linear(x, w[:, :, 0], sd[key + ".conv.bias"])


def bytenet_block(x: Tensor, sd, pfx: str, dilation: int, act: str) -> Tensor:
    """sequence_models ByteNetBlock on (N, L, C) input (restated, see header)."""
    a = _act(act)
    h = a(_ln(x, sd, pfx + ".sequence1.0"))
    h = _pff(h, sd, pfx + ".sequence1.2")
    h = a(_ln(h, sd, pfx + ".sequence1.3"))
    w = sd[pfx + ".conv.weight"]          # [d_h, d_h, k]
    k = w.shape[-1]
    h = F.conv1d(h.transpose(1, 2), w, sd[pfx + ".conv.bias"],
                 padding=dilation * (k - 1) // 2, dilation=dilation).transpose(1, 2)
    h = a(_ln(h, sd, pfx + ".sequence2.0"))
    h = _pff(h, sd, pfx + ".sequence2.2")
    return x + h


def _dilations(n_layers: int, r: int) -> List[int]:
    # model/encoder/model.py:134-135, 256-257
    log2 = int(math.log2(r)) + 1
    return [2 ** (n % log2) for n in range(n_layers)]


def rope_rotate(x: Tensor, rope: Tensor) -> Tensor:
    """cross_attention.py:59-88 -- adjacent pairs (2j, 2j+1) as complex numbers.

    x: [B, L, H, D]; rope: complex64 [L, D/2] (the state-dict buffer).
    """
    xr = x[..., 0::2]
    xi = x[..., 1::2]
    c = rope.real[None, :, None, :]
    s = rope.imag[None, :, None, :]
    out = torch.empty_like(x)
    out[..., 0::2] = xr * c - xi * s
    out[..., 1::2] = xr * s + xi * c
    return out


def att_layer(x: Tensor, sd, pfx: str, nhead: int) -> Tensor:
    """cross_attention.py:131-173 (self-attention, no mask, pads attended)."""
    B, L, _ = x.shape
    q = F.linear(x, sd[pfx + ".query.weight"], sd[pfx + ".query.bias"])
    k = F.linear(x, sd[pfx + ".key.weight"], sd[pfx + ".key.bias"])
    v = F.linear(x, sd[pfx + ".value.weight"], sd[pfx + ".value.bias"])
    att_model = q.shape[-1]
    dk = (att_model / nhead) ** 0.5                       # :136
    q = q.view(B, L, nhead, -1)
    k = k.view(B, L, nhead, -1)
    v = v.view(B, L, nhead, -1)
    rope = sd[pfx + ".rope"]
    assert rope.shape[0] == L, "RoPE length must equal sequence length (cross_attention.py:30)"
    q = rope_rotate(q, rope).transpose(1, 2)
    k = rope_rotate(k, rope).transpose(1, 2)
    v = v.transpose(1, 2)
    w = torch.softmax(torch.matmul(q, k.transpose(-2, -1)) / dk, dim=-1)   # :168-169
    o = torch.matmul(w, v).permute(0, 2, 1, 3).reshape(B, L, att_model)
    return F.linear(o, sd[pfx + ".out_put.weight"], sd[pfx + ".out_put.bias"])


def self_att_block(x: Tensor, sd, pfx: str, nhead: int) -> Tensor:
    """cross_attention.py:273-287; the FFN residual is the block INPUT."""
    a = x + att_layer(x, sd, pfx + ".attn_hl", nhead)
    a = a + att_layer(_ln(a, sd, pfx + ".norm_hl1"), sd, pfx + ".attn_hl_c", nhead)
    f = _ln(a, sd, pfx + ".norm_hl2")
    f = F.relu(F.linear(f, sd[pfx + ".ff_hl.0.weight"], sd[pfx + ".ff_hl.0.bias"]))
    f = F.linear(f, sd[pfx + ".ff_hl.2.weight"], sd[pfx + ".ff_hl.2.bias"])
    return f + x


def region_embed(region: Tensor, sd) -> Tensor:
    """model.py:208-230 RegionEmbedder (Nb: nanoencoder/model.py:192-213)."""
    x = F.embedding(region, sd["region_encoder.region_embedding.weight"])
    x = F.relu(_ln(x, sd, "region_encoder.region_layer1.0"))
    x = _pff(x, sd, "region_encoder.region_layer1.2")
    return F.relu(_ln(x, sd, "region_encoder.region_layer1.3"))


def pos_embed(r: Tensor, sd) -> Tensor:
    """model.py:233-246 PosEmbedder; MLP.ln1/ln2 are Linear layers (model.py:23-25)."""
    L = r.shape[1]
    pe = sd["pos_encoder.pos_embedding.pe"][:L]           # [L,1,256]
    x = r + pe.reshape(1, L, -1)
    m = F.linear(x, sd["pos_encoder.pos_lin.ln1.weight"], sd["pos_encoder.pos_lin.ln1.bias"])
    m = F.linear(F.gelu(m), sd["pos_encoder.pos_lin.ln2.weight"], sd["pos_encoder.pos_lin.ln2.bias"])
    return x + m                                          # MLP.dropout is a Module -> off in eval


def side_embed(chain: Tensor, sd, B: int) -> Tensor:
    """model.py:183-205 SideEmbedder: rows with side==0 tile x152, side!=0 tile x139."""
    e = F.embedding(chain.view(-1, 1), sd["side_encoder.side_embeddinng.weight"])
    e = F.linear(e, sd["side_encoder.side_mlp.0.weight"], sd["side_encoder.side_mlp.0.bias"])
    e = F.relu(F.layer_norm(e, (e.shape[-1],), sd["side_encoder.side_mlp.1.weight"],
                            sd["side_encoder.side_mlp.1.bias"], 1e-5))
    e = F.linear(e, sd["side_encoder.side_mlp.3.weight"], sd["side_encoder.side_mlp.3.bias"])
    h = e[chain == 0].repeat(1, H_LEN, 1)
    l = e[chain != 0].repeat(1, L_LEN, 1)
    assert h.shape[0] == B and l.shape[0] == B, "chain must be [0]*B + [1|2]*B (sample.py:174-175)"
    return torch.cat((h, l), dim=1)


def _head(x: Tensor, sd) -> Tensor:
    return F.linear(_ln(x, sd, "last_norm"), sd["decoder.weight"], sd["decoder.bias"])


# --------------------------------------------------------------------------- #
# full forwards                                                                #
# --------------------------------------------------------------------------- #
def forward_ab(sd: Dict[str, Tensor], cfg: dict, tokens: Tensor, region: Tensor,
               chain: Tensor, drop=None, trace: Optional[dict] = None) -> Tensor:
    """AntiTFNet.forward, model/encoder/model.py:351-384.  Returns [B, 291, 23].

    `trace` (optional dict) receives intermediate activations for kernel localisation."""
    drop = drop or DropoutOff()
    B = tokens.shape[0]
    p = float(cfg.get("dropout", 0.0))
    dil = _dilations(cfg["n_encoder_layers"], cfg["r"])
    act = cfg.get("activation", "relu")
    e = F.embedding(tokens.long(), sd["aa_encoder.embedder.weight"])
    h_e, l_e = e[:, :H_LEN], e[:, H_LEN:]
    for n, d in enumerate(dil):                            # model.py:170-180
        h_e = bytenet_block(h_e, sd, f"aa_encoder.h_layers.{n}", d, act)
        l_e = bytenet_block(l_e, sd, f"aa_encoder.l_layers.{n}", d, act)
        if p > 0.0:
            h_e = _apply_dropout(h_e, p, drop)
            l_e = _apply_dropout(l_e, p, drop)
    e = torch.cat((h_e, l_e), dim=1)
    if trace is not None:
        trace["E"] = e
    c = side_embed(chain.long(), sd, B)
    pos = pos_embed(region_embed(region.long(), sd), sd)
    e = e + pos + c                                        # model.py:357
    x = torch.cat((e, pos, c), dim=-1)                     # model.py:358
    if trace is not None:
        trace["X0"] = x
    h_s, l_s = x[:, :H_LEN], x[:, H_LEN:]
    dil2 = _dilations(cfg["dual_layers"], cfg["r"])
    # DualConv: activation NOT forwarded -> 'relu' (model.py:345,252); p = F.dropout default 0.5
    for n, d in enumerate(dil2):                           # model.py:292-297
        h_s = bytenet_block(h_s, sd, f"dual_conv_block.h_layers.{n}", d, "relu")
        if p > 0.0:
            h_s = _apply_dropout(h_s, 0.5, drop)
    for n, d in enumerate(dil2):                           # model.py:299-304
        l_s = bytenet_block(l_s, sd, f"dual_conv_block.l_layers.{n}", d, "relu")
        if p > 0.0:
            l_s = _apply_dropout(l_s, 0.5, drop)
    x = torch.cat((h_s, l_s), dim=1)
    if trace is not None:
        trace["Xconv"] = x
    for n in range(cfg["cs_layers"]):
        x = self_att_block(x, sd, f"self_at.layers.{n}", cfg["nhead"])
    if trace is not None:
        trace["X"] = x
    return _head(x, sd)


def forward_nb(sd: Dict[str, Tensor], cfg: dict, tokens: Tensor, region: Tensor,
               chain=None, drop=None, trace: Optional[dict] = None) -> Tensor:
    """NanoAntiTFNet.forward, model/nanoencoder/model.py:313-343.  [B, 152, 23]."""
    drop = drop or DropoutOff()
    p = float(cfg.get("dropout", 0.0))
    act = cfg.get("activation", "relu")
    e = F.embedding(tokens.long(), sd["aa_encoder.embedder.weight"])
    for n, d in enumerate(_dilations(cfg["n_encoder_layers"], cfg["r"])):   # :165-170
        e = bytenet_block(e, sd, f"aa_encoder.layers.{n}", d, act)
        if p > 0.0:
            e = _apply_dropout(e, p, drop)
    if trace is not None:
        trace["E"] = e
    pos = pos_embed(region_embed(region.long(), sd), sd)
    e = e + pos
    x = torch.cat((e, pos), dim=-1)                        # :317-318
    if trace is not None:
        trace["X0"] = x
    # NanoConv default activation is 'gelu' (nanoencoder/model.py:242), p = 0.5 default
    for n, d in enumerate(_dilations(cfg["dual_layers"], cfg["r"])):        # :261-266
        x = bytenet_block(x, sd, f"nano_conv_block.layers.{n}", d, "gelu")
        if p > 0.0:
            x = _apply_dropout(x, 0.5, drop)
    if trace is not None:
        trace["Xconv"] = x
    for n in range(cfg["cs_layers"]):
        x = self_att_block(x, sd, f"self_at.layers.{n}", cfg["nhead"])
    if trace is not None:
        trace["X"] = x
    return _head(x, sd)


def forward(kind: str, sd, cfg, tokens, region, chain=None, drop=None, trace=None) -> Tensor:
    if kind == "ab":
        return forward_ab(sd, cfg, tokens, region, chain, drop, trace)
    return forward_nb(sd, cfg, tokens, region, chain, drop, trace)


# --------------------------------------------------------------------------- #
# sampling loop  (sample.py:497-513, nanosample.py:314-329)                     #
# --------------------------------------------------------------------------- #
def sample_step(logits_rows: Tensor, q: Tensor) -> Tensor:
    """softmax over the first 22 classes then torch.multinomial(.,1) == argmax(p/q)."""
    p = torch.softmax(logits_rows[:, :N_SAMPLED_TOKENS], dim=1)
    return torch.argmax(p / q, dim=-1)


def sample_loop(kind: str, sd, cfg, tokens: Tensor, region: Tensor, chain: Optional[Tensor],
                order: Tensor, stream=None, return_logits: bool = False):
    """Order-agnostic denoising loop.

    order: int [T] (reference: one shuffled `loc` shared by all replicas) or
    int [B, T] (per-sequence positions -- the generalisation the B200 path adds;
    the reference is the all-rows-equal special case).
    stream: TorchStream (bit-identical to the reference under the same seed) or
    Recorded.  Per step the stream is consumed as: dropout masks in forward
    order, then one Exp(1) draw of shape [B, 22].
    """
    stream = stream or TorchStream()
    tokens = tokens.clone()
    B = tokens.shape[0]
    order = torch.as_tensor(order, dtype=torch.long)
    if order.dim() == 1:
        order = order.unsqueeze(0).expand(B, -1)
    rows = torch.arange(B)
    logs = []
    for t in range(order.shape[1]):
        logits = forward(kind, sd, cfg, tokens, region, chain, stream)
        pos = order[:, t]
        sel = logits[rows, pos]                            # [B, 23]
        q = stream.exponential((B, N_SAMPLED_TOKENS))
        tokens[rows, pos] = sample_step(sel, q)
        if return_logits:
            logs.append(sel.clone())
    if return_logits:
        return tokens, torch.stack(logs, 0)
    return tokens
