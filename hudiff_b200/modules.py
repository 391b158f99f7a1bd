"""Drop-in model objects for the HuDiff sampling path.

`AntiTFNetB200` / `NanoAntiTFNetB200` take the same constructor kwargs as the
reference's `AntiTFNet` (model/encoder/model.py:325-348) and `NanoAntiTFNet`
(model/nanoencoder/model.py:290-311), expose the same `state_dict()` key set
(so `load_state_dict(ckpt['model'])` works unchanged, including the `pe` and
complex64 `rope` buffers) and consume the torch RNG identically at construction
(sub-modules are created in the reference's order with the same nn classes), so
`seed_all(s); model_selected(cfg)` leaves the generator in the same state.

They hold parameters only.  All arithmetic of `forward` / `sample` runs in the
sm_100a CUDA library behind the C-ABI (`include/hudiff.h`); there is no PyTorch
or CPU fallback -- calling them without the built library or on a CPU tensor
raises.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from .tables import H_LEN, L_LEN


# --------------------------------------------------------------------------- #
# parameter containers (names/shapes/init order mirror the reference)           #
# --------------------------------------------------------------------------- #
class _Act(nn.Module):
    """Parameter-free placeholder keeping nn.Sequential indices aligned."""

    def __init__(self, name):
        super().__init__()
        self.name = name


class _PFF(nn.Module):            # sequence_models PositionFeedForward: attribute `.conv`
    def __init__(self, d_in, d_out):
        super().__init__()
        self.conv = nn.Conv1d(d_in, d_out, 1)


class _ByteNetParams(nn.Module):  # sequence_models ByteNetBlock parameter layout
    def __init__(self, d_in, d_h, d_out, kernel_size, dilation, activation):
        super().__init__()
        self.conv = nn.Conv1d(d_h, d_h, kernel_size, dilation=dilation,
                              padding=dilation * (kernel_size - 1) // 2)
        self.sequence1 = nn.Sequential(nn.LayerNorm(d_in), _Act(activation), _PFF(d_in, d_h),
                                       nn.LayerNorm(d_h), _Act(activation))
        self.sequence2 = nn.Sequential(nn.LayerNorm(d_h), _Act(activation), _PFF(d_h, d_out))
        self.dilation = dilation


def _dilations(n_layers, r):
    log2 = int(np.log2(r)) + 1
    return [2 ** (n % log2) for n in range(n_layers)]


def _stack(d_model, n_layers, k, r, activation):
    return nn.ModuleList([_ByteNetParams(d_model, d_model // 2, d_model, k, d, activation)
                          for d in _dilations(n_layers, r)])


class _AAEncoderAb(nn.Module):    # ByteNetTime, model.py:90-153
    def __init__(self, n_tokens, d_embedding, d_model, n_layers, k, r, activation):
        super().__init__()
        assert d_model == d_embedding
        self.embedder = nn.Embedding(n_tokens, d_embedding)
        self.h_layers = _stack(d_model, n_layers, k, r, activation)
        self.l_layers = _stack(d_model, n_layers, k, r, activation)


class _AAEncoderNb(nn.Module):    # NanoByteNetTime, nanoencoder/model.py:91-148
    def __init__(self, n_tokens, d_embedding, d_model, n_layers, k, r, activation):
        super().__init__()
        assert d_model == d_embedding
        self.embedder = nn.Embedding(n_tokens, d_embedding)
        self.layers = _stack(d_model, n_layers, k, r, activation)


class _Side(nn.Module):           # SideEmbedder, model.py:183-195 (attribute name typo kept)
    def __init__(self, n_side, s_embedding, d_side):
        super().__init__()
        self.side_embeddinng = nn.Embedding(n_side, s_embedding)
        self.side_mlp = nn.Sequential(nn.Linear(s_embedding, d_side), nn.LayerNorm(d_side),
                                      _Act("relu"), nn.Linear(d_side, d_side))


class _Region(nn.Module):         # RegionEmbedder, model.py:208-219
    def __init__(self, n_region, r_embedding, r_model):
        super().__init__()
        self.region_embedding = nn.Embedding(n_region, r_embedding)
        self.region_layer1 = nn.Sequential(nn.LayerNorm(r_embedding), _Act("relu"),
                                           _PFF(r_embedding, r_model), nn.LayerNorm(r_model),
                                           _Act("relu"))


class _SinCos(nn.Module):         # PositionalEncoding, model.py:62-78
    def __init__(self, d_model, max_len):
        super().__init__()
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-np.log(10000.0) / d_model))
        pe = torch.zeros(max_len, 1, d_model)
        pe[:, 0, 0::2] = torch.sin(position * div_term)
        pe[:, 0, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)


class _PosMLP(nn.Module):         # MLP, model.py:19-26 ("ln1"/"ln2" are Linear layers)
    def __init__(self, n):
        super().__init__()
        self.ln1 = nn.Linear(n, 2 * n)
        self.ln2 = nn.Linear(2 * n, n)


class _Pos(nn.Module):            # PosEmbedder, model.py:233-240
    def __init__(self, p_emb, max_len):
        super().__init__()
        self.pos_embedding = _SinCos(p_emb, max_len)
        self.pos_lin = _PosMLP(p_emb)


class _DualStacks(nn.Module):     # DualConv, model.py:249-275 (activation stays 'relu')
    def __init__(self, d_model, n_layers, k, r):
        super().__init__()
        self.h_layers = _stack(d_model, n_layers, k, r, "relu")
        self.l_layers = _stack(d_model, n_layers, k, r, "relu")


class _NanoStack(nn.Module):      # NanoConv, nanoencoder/model.py:239-259 (default 'gelu')
    def __init__(self, d_model, n_layers, k, r):
        super().__init__()
        self.layers = _stack(d_model, n_layers, k, r, "gelu")


class _Att(nn.Module):            # AttLayer, cross_attention.py:131-146
    def __init__(self, d_model, att_model, nhead, length):
        super().__init__()
        self.query = nn.Linear(d_model, att_model)
        self.key = nn.Linear(d_model, att_model)
        self.value = nn.Linear(d_model, att_model)
        self.out_put = nn.Linear(att_model, d_model)
        dim = att_model // nhead
        freqs = 1.0 / (10000.0 ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
        ang = torch.outer(torch.arange(length), freqs).float()
        self.register_buffer("rope", torch.polar(torch.ones_like(ang), ang))   # complex64 [L, dim/2]


class _AttBlock(nn.Module):       # SelfAttBlock, cross_attention.py:257-271
    def __init__(self, d_model, att_model, dim_ff, nhead, length):
        super().__init__()
        self.attn_hl = _Att(d_model, att_model, nhead, length)
        self.attn_hl_c = _Att(d_model, att_model, nhead, length)
        self.norm_hl1 = nn.LayerNorm(d_model)
        self.norm_hl2 = nn.LayerNorm(d_model)
        self.ff_hl = nn.Sequential(nn.Linear(d_model, dim_ff), _Act("relu"), nn.Linear(dim_ff, d_model))


class _AttNet(nn.Module):         # SelfAttNet, cross_attention.py:291-297
    def __init__(self, d_model, att_model, dim_ff, nhead, length, n):
        super().__init__()
        self.layers = nn.ModuleList([_AttBlock(d_model, att_model, dim_ff, nhead, length)
                                     for _ in range(n)])


# --------------------------------------------------------------------------- #
# the two denoisers                                                            #
# --------------------------------------------------------------------------- #
class _B200Denoiser(nn.Module):
    """Shared device plumbing: lazily (re)packs weights into a native engine."""

    kind = "?"
    seq_len = 0

    def _post_init(self, cfg: dict):
        self.cfg = dict(cfg)
        self._engine = None
        self._engine_key = None
        self.precision = "fp16"          # 'fp16' (TF32-grade operands) | 'fp16x3' (split, ~fp32)

    # -- engine management -------------------------------------------------- #
    def _weights_key(self, device):
        p = next(self.parameters())
        ver = sum(int(t._version) for t in self.state_dict().values())
        return (str(device), p.data_ptr(), ver, self.precision)

    def engine(self, device=None):
        from .engine import Engine       # imported lazily: needs the built CUDA library
        device = torch.device(device) if device is not None else next(self.parameters()).device
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if device.type != "cuda":
            raise RuntimeError(
                "hudiff_b200 models only run on CUDA devices (sm_100a); there is no CPU path. "
                "Move the model with .to('cuda').")
        key = self._weights_key(device)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(self.kind, self.cfg, self.state_dict(), device, self.precision)
            self._engine_key = key
        return self._engine

    def set_precision(self, precision: str):
        assert precision in ("fp16", "fp16x3")
        self.precision = precision
        return self

    # -- reference surface ---------------------------------------------------- #
    def _check_inputs(self, tokens, region):
        if tokens.dim() != 2 or tokens.shape[1] != self.seq_len:
            # the reference asserts this inside RoPE (cross_attention.py:30)
            raise AssertionError(
                f"sequence length {tuple(tokens.shape)} must be [B, {self.seq_len}] (max_len)")
        assert region.shape == tokens.shape

    @torch.no_grad()
    def sample(self, tokens, region, chain=None, order=None, *, rng="philox", seed=0,
               dropout="reference", return_logits=False, seq_offset=0, check=True):
        """Run the whole T-step denoising loop on the device.

        tokens [B, L] int (masked slots = 22), region [B, L], chain [2B] (Ab) or None,
        order int [T] or [B, T]: slot sampled at step t for each sequence.
        rng: 'philox' (in-kernel counter RNG, throughput mode) or 'torch_cpu_replay'
        (host replays torch's CPU generator stream -> token-for-token equal to the
        reference loop sample.py:499-513 under the same seed; needs B small).
        dropout: 'reference' (live eval-mode dropout as the reference) or 'off'.
        seed / seq_offset: Philox key and the global index of this batch's first sequence; shards of one
        batch that share a seed pass their offset (sequence b draws the stream of sequence seq_offset + b),
        and repeated calls must change one of the two or they repeat the same draws.
        Returns the completed tokens [B, L] (same dtype/device as `tokens`).
        """
        self._check_inputs(tokens, region)
        return self.engine(tokens.device).sample(tokens, region, chain, order, rng=rng, seed=seed,
                                                 dropout=dropout, return_logits=return_logits,
                                                 seq_offset=seq_offset, check=check)


class AntiTFNetB200(_B200Denoiser):
    """HuDiff-Ab denoiser (paired heavy 152 + light 139 IMGT slots = 291)."""

    kind = "ab"
    seq_len = H_LEN + L_LEN

    def __init__(self, n_tokens, d_embedding, d_model, n_encoder_layers, aa_kernel_size, r,
                 n_side, s_embedding, s_model, n_region, r_embedding, r_model, n_pos_model, max_len,
                 sum_d_model, dual_layers, att_model, dim_feedforward, nhead, cs_layers,
                 rank=None, n_frozen_embs=None, padding_idx=None, causal=False, dropout=0.0,
                 slim=True, activation="relu", down_embed=False, timesteps=None):
        super().__init__()
        if rank is not None or n_frozen_embs is not None or padding_idx is not None or causal \
                or not slim or down_embed:
            raise NotImplementedError("only the option set used by the released configs is supported")
        assert max_len == self.seq_len, "HuDiff-Ab is hard-wired to 152+139 slots (model.py:185)"
        assert sum_d_model == d_model + n_pos_model + s_model and r_model == n_pos_model == d_model
        self.aa_encoder = _AAEncoderAb(n_tokens, d_embedding, d_model, n_encoder_layers,
                                       aa_kernel_size, r, activation)
        self.side_encoder = _Side(n_side, s_embedding, s_model)
        self.region_encoder = _Region(n_region, r_embedding, r_model)
        self.pos_encoder = _Pos(n_pos_model, max_len)
        self.dual_conv_block = _DualStacks(sum_d_model, dual_layers, aa_kernel_size, r)
        self.self_at = _AttNet(sum_d_model, att_model, dim_feedforward, nhead, max_len, cs_layers)
        self.last_norm = nn.LayerNorm(sum_d_model)
        self.decoder = nn.Linear(sum_d_model, n_tokens)
        self._post_init(dict(
            n_tokens=n_tokens, d_model=d_model, n_encoder_layers=n_encoder_layers,
            aa_kernel_size=aa_kernel_size, r=r, n_side=n_side, s_embedding=s_embedding,
            n_region=n_region, r_embedding=r_embedding, max_len=max_len, sum_d_model=sum_d_model,
            dual_layers=dual_layers, att_model=att_model, dim_feedforward=dim_feedforward,
            nhead=nhead, cs_layers=cs_layers, dropout=float(dropout), activation=activation))

    @torch.no_grad()
    def forward(self, H_L_seq, H_L_region_type, H_L_chn_type, *, dropout="off", seed=0):
        """Same call as the reference forward (model.py:366-384): -> fp32 [B, 291, 23].

        `dropout='off'` gives deterministic logits; `dropout='torch_cpu_replay'`
        reproduces the reference's live eval-mode dropout from torch's CPU stream.
        """
        self._check_inputs(H_L_seq, H_L_region_type)
        return self.engine(H_L_seq.device).forward(H_L_seq, H_L_region_type, H_L_chn_type,
                                                   dropout=dropout, seed=seed)


class NanoAntiTFNetB200(_B200Denoiser):
    """HuDiff-Nb denoiser (single VHH chain, 152 IMGT slots)."""

    kind = "nb"
    seq_len = H_LEN

    def __init__(self, n_tokens, d_embedding, d_model, n_encoder_layers, aa_kernel_size, r,
                 n_region, r_embedding, r_model, n_pos_model, max_len, sum_d_model, dual_layers,
                 att_model, dim_feedforward, nhead, cs_layers,
                 rank=None, n_frozen_embs=None, padding_idx=None, causal=False, dropout=0.0,
                 slim=True, activation="relu", down_embed=False, timesteps=None):
        super().__init__()
        if rank is not None or n_frozen_embs is not None or padding_idx is not None or causal \
                or not slim or down_embed:
            raise NotImplementedError("only the option set used by the released configs is supported")
        assert max_len == self.seq_len
        assert sum_d_model == d_model + n_pos_model and r_model == n_pos_model == d_model
        self.aa_encoder = _AAEncoderNb(n_tokens, d_embedding, d_model, n_encoder_layers,
                                       aa_kernel_size, r, activation)
        self.region_encoder = _Region(n_region, r_embedding, r_model)
        self.pos_encoder = _Pos(n_pos_model, max_len)
        self.nano_conv_block = _NanoStack(sum_d_model, dual_layers, aa_kernel_size, r)
        self.self_at = _AttNet(sum_d_model, att_model, dim_feedforward, nhead, max_len, cs_layers)
        self.last_norm = nn.LayerNorm(sum_d_model)
        self.decoder = nn.Linear(sum_d_model, n_tokens)
        self._post_init(dict(
            n_tokens=n_tokens, d_model=d_model, n_encoder_layers=n_encoder_layers,
            aa_kernel_size=aa_kernel_size, r=r, n_region=n_region, r_embedding=r_embedding,
            max_len=max_len, sum_d_model=sum_d_model, dual_layers=dual_layers, att_model=att_model,
            dim_feedforward=dim_feedforward, nhead=nhead, cs_layers=cs_layers,
            dropout=float(dropout), activation=activation))

    @torch.no_grad()
    def forward(self, H_seq, H_region_type, H_chn_type=None, *, dropout="off", seed=0):
        """Same call as nanoencoder/model.py:325-343 (chain type accepted and ignored)."""
        self._check_inputs(H_seq, H_region_type)
        return self.engine(H_seq.device).forward(H_seq, H_region_type, None, dropout=dropout, seed=seed)


def model_selected(config, pretrained_model=None, tokenizer=None):
    """utils/train_utils.py:43-55 for the two sampling-path model names."""
    name = config["name"] if isinstance(config, dict) else config.name
    model_cfg = config["model"] if isinstance(config, dict) else config.model
    if name == "trans_oadm":
        return AntiTFNetB200(**model_cfg)
    if name == "nano":
        return NanoAntiTFNetB200(**model_cfg)
    raise NotImplementedError(
        f"config.name={name!r}: only 'trans_oadm' and 'nano' are on the sampling path "
        "(fine-tune frameworks are training-only, SURVEY.md §2 rows 14-15)")
