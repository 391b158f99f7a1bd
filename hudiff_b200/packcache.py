"""`.hudiffpack`: a flat, hash-keyed cache of a reference checkpoint.

The reference's `.pt` files are Python pickles of an optimizer-era training state (model weights under one of
four layouts, an `easydict` config, optimizer moments, ...; sample.py:446-458, nanosample.py:283-295,
nanofinetune.py:531-539).  Every CLI start un-pickles all of it to use ~25 % of it.  The pack keeps exactly
what the sampling path consumes -- the resolved config and the resolved state dict -- in a layout that is read
with ONE `readinto` and no Python object construction per tensor:

    magic "HUDIFFPK1\n" | u64 header length | header JSON | padding to 64 B | tensor bytes (64-byte aligned)

The header records the key (sha256 of the checkpoint bytes + how it was resolved), the config as plain JSON and
one entry per tensor (name, dtype, shape, offset, nbytes).  A pack whose key does not match the checkpoint on
disk, or that fails any structural check, is ignored and rebuilt; writing is atomic (temp file + rename) and a
read-only checkpoint directory just disables the cache.  Nothing here touches the GPU: the packed weights are
still laid out for the tensor cores by `hudiff_load_weights` (milliseconds, on the device).
"""
from __future__ import annotations

import hashlib
import json
import os
import struct
import tempfile
from typing import Dict, Optional, Tuple

import torch

from .config import EasyDict

MAGIC = b"HUDIFFPK1\n"
ALIGN = 64
_DTYPES = {"float32": torch.float32, "float64": torch.float64, "float16": torch.float16, "bfloat16": torch.bfloat16,
           "int64": torch.int64, "int32": torch.int32, "int8": torch.int8, "uint8": torch.uint8, "bool": torch.bool,
           "complex64": torch.complex64}
_NAMES = {v: k for k, v in _DTYPES.items()}


def checkpoint_key(path: str, version: str, kind: str) -> str:
    """sha256 over the checkpoint bytes and the way it is resolved (layout version, model family)."""
    h = hashlib.sha256()
    h.update(f"hudiffpack-1|{version}|{kind}|".encode())
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def pack_path_for(path: str) -> str:
    """`<ckpt>.hudiffpack` next to the checkpoint, or under $HUDIFF_PACK_DIR (flat, by file name) when set."""
    d = os.environ.get("HUDIFF_PACK_DIR")
    if d:
        return os.path.join(d, os.path.basename(path) + ".hudiffpack")
    return path + ".hudiffpack"


def _plain(x):
    """Config -> JSON-able (EasyDict / dict / list / tuple / scalars; anything else as its repr, marked)."""
    if isinstance(x, dict):
        return {str(k): _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    if isinstance(x, (str, int, float, bool)) or x is None:
        return x
    if isinstance(x, torch.Tensor) and x.numel() == 1:
        return x.item()
    return {"__repr__": repr(x)}


def _rich(x):
    if isinstance(x, dict):
        return EasyDict({k: _rich(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_rich(v) for v in x]
    return x


def write_pack(pack_path: str, key: str, config, state: Dict[str, torch.Tensor]) -> bool:
    """Serialise (config, state); returns False (and leaves nothing behind) if the location is not writable."""
    entries, off = [], 0
    blobs = []
    for name, t in state.items():
        if not isinstance(t, torch.Tensor) or t.dtype not in _NAMES:
            return False                                   # unexpected content: do not cache what cannot be restored exactly
        c = t.detach().cpu().contiguous()
        raw = torch.view_as_real(c) if c.is_complex() else c
        b = raw.reshape(-1).view(torch.uint8).numpy().tobytes() if raw.numel() else b""
        off = (off + ALIGN - 1) // ALIGN * ALIGN
        entries.append({"name": name, "dtype": _NAMES[t.dtype], "shape": list(t.shape), "offset": off, "nbytes": len(b)})
        blobs.append((off, b))
        off += len(b)
    header = json.dumps({"key": key, "config": _plain(config), "tensors": entries, "data_bytes": off}).encode()
    head_len = len(MAGIC) + 8 + len(header)
    data0 = (head_len + ALIGN - 1) // ALIGN * ALIGN
    try:
        fd, tmp = tempfile.mkstemp(prefix=".hudiffpack.", dir=os.path.dirname(os.path.abspath(pack_path)) or ".")
    except OSError:
        return False
    try:
        with os.fdopen(fd, "wb") as f:
            f.write(MAGIC)
            f.write(struct.pack("<Q", len(header)))
            f.write(header)
            f.write(b"\0" * (data0 - head_len))
            pos = 0
            for o, b in blobs:
                if o > pos:
                    f.write(b"\0" * (o - pos))
                f.write(b)
                pos = o + len(b)
        os.replace(tmp, pack_path)
        return True
    except OSError:
        try:
            os.unlink(tmp)
        except OSError:
            pass
        return False


def read_pack(pack_path: str, key: str) -> Optional[Tuple[EasyDict, Dict[str, torch.Tensor]]]:
    """(config, state) if the pack exists, is well formed and was built from the checkpoint `key` names; else None."""
    try:
        with open(pack_path, "rb") as f:
            if f.read(len(MAGIC)) != MAGIC:
                return None
            (hl,) = struct.unpack("<Q", f.read(8))
            if hl > (64 << 20):
                return None
            header = json.loads(f.read(hl).decode())
            if header.get("key") != key:
                return None
            head_len = len(MAGIC) + 8 + hl
            f.seek((head_len + ALIGN - 1) // ALIGN * ALIGN)
            n = int(header["data_bytes"])
            buf = bytearray(n)
            if f.readinto(buf) != n:
                return None
    except (OSError, ValueError, KeyError, struct.error):
        return None
    state: Dict[str, torch.Tensor] = {}
    try:
        for e in header["tensors"]:
            dt = _DTYPES[e["dtype"]]
            shape, off, nb = tuple(e["shape"]), int(e["offset"]), int(e["nbytes"])
            if off + nb > n:
                return None
            if nb == 0:
                state[e["name"]] = torch.empty(shape, dtype=dt)
                continue
            if dt == torch.complex64:
                t = torch.view_as_complex(torch.frombuffer(buf, dtype=torch.float32, count=nb // 4, offset=off).view(*shape, 2))
            else:
                t = torch.frombuffer(buf, dtype=dt, count=nb // dt.itemsize, offset=off).view(shape)
            state[e["name"]] = t
    except (KeyError, RuntimeError, ValueError):
        return None
    return _rich(header["config"]), state


def load_cached(path: str, version: str, kind: str, resolve):
    """(config, state) of checkpoint `path`: from its pack when the key matches, else through `resolve(path, version,
    kind)` (the un-pickling loader), after which the pack is (re)written.  HUDIFF_PACK=0 bypasses the cache."""
    if os.environ.get("HUDIFF_PACK", "1") == "0":
        return resolve(path, version, kind)
    key = checkpoint_key(path, version, kind)
    pp = pack_path_for(path)
    hit = read_pack(pp, key)
    if hit is not None:
        return hit
    config, state = resolve(path, version, kind)
    write_pack(pp, key, config, state)
    return config, state
