"""Python owner of one native engine handle (weights packed for one device).

Thin plumbing over the C-ABI: converts the reference-shaped state_dict into `hudiff_tensor`
records, moves integer inputs to int32 device tensors, hands raw device pointers and the current
CUDA stream to the library.  No model arithmetic happens here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi, rng as rng_mod


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


class Engine:
    def __init__(self, kind: str, cfg: dict, state_dict, device, precision: str = "fp16"):
        self.lib = capi.load()
        self.kind, self.cfg = kind, dict(cfg)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise capi.HudiffError("hudiff_b200 engines live on CUDA devices only")
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.L = cfg["max_len"]
        self.n_tokens = cfg["n_tokens"]
        c = capi.HudiffCfg(
            kind=capi.HUDIFF_AB if kind == "ab" else capi.HUDIFF_NB, n_tokens=cfg["n_tokens"],
            d_model=cfg["d_model"], n_encoder_layers=cfg["n_encoder_layers"],
            aa_kernel_size=cfg["aa_kernel_size"], r=cfg["r"], n_side=cfg.get("n_side", 0),
            s_embedding=cfg.get("s_embedding", 0), n_region=cfg["n_region"], r_embedding=cfg["r_embedding"],
            max_len=cfg["max_len"], sum_d_model=cfg["sum_d_model"], dual_layers=cfg["dual_layers"],
            att_model=cfg["att_model"], dim_feedforward=cfg["dim_feedforward"], nhead=cfg["nhead"],
            cs_layers=cfg["cs_layers"], activation=capi.ACT[cfg.get("activation", "relu")],
            precision=capi.HUDIFF_FP16X3 if precision == "fp16x3" else capi.HUDIFF_FP16,
            dropout=float(cfg.get("dropout", 0.0)))
        h = C.c_void_p()
        capi.check(self.lib.hudiff_create(C.byref(h), C.byref(c), self.dev_index), "hudiff_create")
        self.h = h
        keep, recs = [], []
        for name, v in state_dict.items():
            v = v.detach()
            dtype = 0
            if v.is_complex():
                v, dtype = torch.view_as_real(v.to(torch.complex64)), 1
            v = v.to(torch.float32).contiguous()
            shape = list(v.shape[:-1]) if dtype == 1 else list(v.shape)
            shape = (shape + [1, 1, 1, 1])[:4] if len(shape) <= 4 else [int(np.prod(shape)), 1, 1, 1]
            keep.append((v, name.encode()))
            recs.append(capi.HudiffTensor(name=keep[-1][1], data=v.data_ptr(), dtype=dtype,
                                          ndim=min(len(v.shape) - (1 if dtype else 0), 4),
                                          shape=(C.c_int64 * 4)(*shape), on_device=1 if v.is_cuda else 0))
        arr = (capi.HudiffTensor * len(recs))(*recs)
        capi.check(self.lib.hudiff_load_weights(self.h, arr, len(recs)), "hudiff_load_weights")
        del keep

    # ------------------------------------------------------------------------------------------ #
    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.hudiff_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_gemm_impl(self, impl: str):
        capi.check(self.lib.hudiff_set_gemm_impl(self.h, 1 if impl == "simt" else 0), "hudiff_set_gemm_impl")

    def last_launch_count(self) -> int:
        return int(self.lib.hudiff_last_launch_count(self.h))

    def workspace_bytes(self, B: int) -> int:
        return int(self.lib.hudiff_workspace_bytes(self.h, B))

    def profile_step(self, tokens, region, chain=None, dropout="reference", seed=0):
        """Per-kernel-class device time of ONE denoiser forward (events around every launch)."""
        B = tokens.shape[0]
        tok, reg, chn = self._i32(tokens), self._i32(region), self._i32(chain)
        live = 1 if (dropout == "reference" and float(self.cfg.get("dropout", 0.0)) > 0.0) else 0
        r = capi.HudiffRng(mode=0, dropout=live, seed=int(seed), dropout_bits=None, exp_draws=None, seq_offset=0)
        ms, n, fl = (C.c_double * 4)(), (C.c_longlong * 4)(), (C.c_double * 4)()
        with torch.cuda.device(self.device):
            capi.check(self.lib.hudiff_profile_step(self.h, _ptr(tok), _ptr(reg), _ptr(chn), B, C.byref(r), ms, n, fl),
                       "hudiff_profile_step")
        names = ("gemm_tc", "attention", "rowwise", "head")
        return {k: {"ms": ms[i], "launches": int(n[i]), "flops": fl[i]} for i, k in enumerate(names)}

    def debug_buffer(self, B: int, name: str) -> torch.Tensor:
        """Copy of a named workspace buffer [B*L, ld] after the last forward (tests only)."""
        ptr, ld, is_half = C.c_void_p(), C.c_int(), C.c_int()
        capi.check(self.lib.hudiff_debug_buffer(self.h, B, name.encode(), C.byref(ptr), C.byref(ld),
                                                C.byref(is_half)), "hudiff_debug_buffer")
        rows = B * self.L
        torch.cuda.synchronize(self.device)

        class _Raw:     # zero-copy view of library-owned memory through the CUDA array interface
            __cuda_array_interface__ = {"shape": (rows, ld.value), "typestr": "<f2" if is_half.value else "<f4",
                                        "data": (ptr.value, False), "version": 2}

        return torch.as_tensor(_Raw(), device=self.device).clone()

    def _i32(self, t):
        return None if t is None else t.to(device=self.device, dtype=torch.int32).contiguous()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------------------------------ #
    def check_error(self):
        """Synchronise the library's stream and raise what the device flagged during the last call: IndexError for an
        out-of-range token / region / chain id or sampling slot, HudiffError for a pipeline watchdog."""
        with torch.cuda.device(self.device):
            capi.check(self.lib.hudiff_check_error(self.h), "hudiff_check_error")

    def forward(self, tokens, region, chain=None, dropout="off", seed=0, seq_offset=0, check=True):
        B = tokens.shape[0]
        tok, reg, chn = self._i32(tokens), self._i32(region), self._i32(chain)
        if self.kind == "ab" and chn is None:
            raise ValueError("HuDiff-Ab forward needs the chain-type vector [h]*B + [l]*B")
        r = capi.HudiffRng(mode=0, dropout=0, seed=int(seed), dropout_bits=None, exp_draws=None, seq_offset=int(seq_offset))
        bits = None
        if dropout == "torch_cpu_replay":
            words = rng_mod.draw_forward_bits(self.kind, self.cfg, B)
            if words.size:
                bits = torch.from_numpy(words.astype(np.uint32).view(np.int32)).to(self.device)
                r.mode, r.dropout, r.dropout_bits = 1, 1, bits.data_ptr()
        elif dropout == "philox":
            r.mode, r.dropout = 0, 1
        elif dropout != "off":
            raise ValueError(f"dropout={dropout!r}")
        out = torch.empty((B, self.L, self.n_tokens), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            capi.check(self.lib.hudiff_forward(self.h, _ptr(tok), _ptr(reg), _ptr(chn), B, C.byref(r),
                                               _ptr(out), self._stream()), "hudiff_forward")
        if check:
            self.check_error()
        elif bits is not None:
            torch.cuda.current_stream(self.device).synchronize()    # keep `bits` alive until consumed
        return out

    def _order(self, order, B):
        order = torch.as_tensor(order)
        if order.dim() == 1:                        # the reference's one shared `loc` (sample.py:498)
            order = order.unsqueeze(0).expand(B, -1)
        if order.shape[0] != B:
            raise ValueError("order must be [T] or [B, T]")
        return order.to(device=self.device, dtype=torch.int32).contiguous()

    def sample(self, tokens, region, chain, order, rng="philox", seed=0, dropout="reference",
               return_logits=False, seq_offset=0, check=True):
        """The whole T-step loop on the device.  `seed` keys the in-kernel Philox streams; sequence b of this call draws
        the stream of global sequence `seq_offset + b`, so shards of one batch (ranks, or successive calls) that share a
        seed must pass their offset -- otherwise they would repeat each other's draws.  `check=True` synchronises and
        raises device-side errors (IndexError for out-of-range ids / slots); `check=False` returns asynchronously."""
        B = tokens.shape[0]
        tok = tokens.to(device=self.device, dtype=torch.int32).contiguous().clone()
        reg, chn = self._i32(region), self._i32(chain)
        if self.kind == "ab" and chn is None:
            raise ValueError("HuDiff-Ab sampling needs the chain-type vector [h]*B + [l]*B")
        order = self._order(order, B)
        T = order.shape[1]
        live = 1 if (dropout == "reference" and float(self.cfg.get("dropout", 0.0)) > 0.0) else 0
        r = capi.HudiffRng(mode=0, dropout=live, seed=int(seed), dropout_bits=None, exp_draws=None, seq_offset=int(seq_offset))
        bits = exps = None
        if rng == "torch_cpu_replay":
            bits, exps = rng_mod.draw_loop_streams(self.kind, self.cfg, B, T, dropout=bool(live))
            exps = exps.to(self.device).contiguous()
            r.mode, r.exp_draws = 1, exps.data_ptr()
            if bits is not None:
                bits = bits.to(self.device).contiguous()
                r.dropout_bits = bits.data_ptr()
        elif rng != "philox":
            raise ValueError(f"rng={rng!r}")
        sel = torch.empty((T, B, self.n_tokens), dtype=torch.float32, device=self.device) if return_logits else None
        with torch.cuda.device(self.device):
            capi.check(self.lib.hudiff_sample(self.h, _ptr(tok), _ptr(reg), _ptr(chn), _ptr(order), B, T,
                                              C.byref(r), _ptr(sel), self._stream()), "hudiff_sample")
        if check:
            self.check_error()
        elif bits is not None or exps is not None:
            torch.cuda.current_stream(self.device).synchronize()
        out = tok.to(dtype=tokens.dtype)
        if tokens.device != self.device:
            out = out.to(tokens.device)
        return (out, sel) if return_logits else out

    def sample_host(self, tokens, region, chain, order, seed=0, dropout="reference", seq_offset=0):
        """End-to-end call with HOST buffers (numpy int32 in, numpy int32 out): hudiff_sample_host."""
        # in/out token buffer in pinned host memory (the caller's array is left untouched); the pinned
        # buffer is cached per shape, page-locking memory on every call costs milliseconds
        src = np.ascontiguousarray(np.asarray(tokens), dtype=np.int32)
        cache = self.__dict__.setdefault("_pinned_tok", {})
        tok_t = cache.get(src.shape)
        if tok_t is None:
            tok_t = cache[src.shape] = torch.empty(src.shape, dtype=torch.int32).pin_memory()
        tok = tok_t.numpy()
        tok[...] = src
        reg = np.ascontiguousarray(np.asarray(region), dtype=np.int32)
        chn = None if chain is None else np.ascontiguousarray(np.asarray(chain), dtype=np.int32)
        order = np.asarray(order)
        B = tok.shape[0]
        if order.ndim == 1:
            order = np.broadcast_to(order[None, :], (B, order.shape[0]))
        order = np.ascontiguousarray(order, dtype=np.int32)
        T = order.shape[1]
        live = 1 if (dropout == "reference" and float(self.cfg.get("dropout", 0.0)) > 0.0) else 0
        r = capi.HudiffRng(mode=0, dropout=live, seed=int(seed), dropout_bits=None, exp_draws=None, seq_offset=int(seq_offset))
        vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None)
        with torch.cuda.device(self.device):
            capi.check(self.lib.hudiff_sample_host(self.h, vp(tok), vp(reg), vp(chn), vp(order), B, T,
                                                   C.byref(r), C.c_void_p(None)), "hudiff_sample_host")
        return tok.copy()
