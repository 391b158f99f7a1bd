"""`.hudiffpack` checkpoint cache (hudiff_b200/packcache.py): the cached (config, state dict) is
bit-identical to what the un-pickling loader resolves for every checkpoint layout of the reference, a stale or damaged
pack is ignored and rebuilt, and an unwritable location only disables the cache."""
import os

import pytest
import torch

from hudiff_b200 import packcache
from hudiff_b200 import sampling as S
from hudiff_b200.config import EasyDict

from test_cli_plumbing import _write_ckpt


def _same(a, b):
    assert set(a) == set(b)
    for k in a:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
        if a[k].is_complex():
            assert torch.equal(torch.view_as_real(a[k]), torch.view_as_real(b[k])), k
        else:
            assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("kind,version", [("ab", "finetune"), ("ab", "pretrain"), ("nb", "finetune_vh"), ("nb", "pretrain")])
def test_second_load_comes_from_the_pack_and_is_bit_identical(tmp_path, monkeypatch, kind, version):
    path, cfg, sd = _write_ckpt(tmp_path, kind, version)
    config1, state1 = S.load_checkpoint(path, version, kind)
    assert os.path.exists(path + ".hudiffpack")
    assert any(t.is_complex() for t in state1.values()), "the RoPE buffers are part of the state dict"

    def boom(*a, **k):
        raise AssertionError("the pickle was opened although a valid pack exists")
    monkeypatch.setattr(torch, "load", boom)
    config2, state2 = S.load_checkpoint(path, version, kind)
    _same(state1, state2)
    _same(state2, sd)
    m1 = config1["model"] if isinstance(config1, dict) else config1.model
    assert dict(config2.model) == dict(m1) == cfg          # attribute access like the reference's EasyDict
    assert isinstance(config2, EasyDict)
    assert config2["name"] == config1["name"]


def test_changed_checkpoint_invalidates_the_pack(tmp_path):
    path, cfg, sd = _write_ckpt(tmp_path, "nb", "pretrain")
    _, s1 = S.load_checkpoint(path, "pretrain", "nb")
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    k0 = next(k for k, v in ckpt["model"].items() if v.dtype == torch.float32)
    ckpt["model"][k0] = ckpt["model"][k0] + 1.0
    torch.save(ckpt, path)
    _, s2 = S.load_checkpoint(path, "pretrain", "nb")
    name = k0[7:] if k0.startswith("module.") else k0
    assert torch.equal(s2[name], s1[name] + 1.0)
    # and the rebuilt pack now serves the new weights
    hit = packcache.read_pack(path + ".hudiffpack", packcache.checkpoint_key(path, "pretrain", "nb"))
    assert hit is not None and torch.equal(hit[1][name], s2[name])
    # a different resolution of the same file (other layout version) has a different key
    assert packcache.read_pack(path + ".hudiffpack", packcache.checkpoint_key(path, "finetune", "nb")) is None


def test_damaged_pack_is_ignored_and_rebuilt(tmp_path):
    path, cfg, sd = _write_ckpt(tmp_path, "ab", "pretrain")
    S.load_checkpoint(path, "pretrain", "ab")
    pp = path + ".hudiffpack"
    size = os.path.getsize(pp)
    with open(pp, "r+b") as f:
        f.truncate(size // 2)
    _, state = S.load_checkpoint(path, "pretrain", "ab")
    _same(state, sd)
    assert os.path.getsize(pp) == size
    with open(pp, "r+b") as f:
        f.write(b"garbage!!!")
    _, state = S.load_checkpoint(path, "pretrain", "ab")
    _same(state, sd)


def test_cache_can_be_bypassed_or_unwritable(tmp_path, monkeypatch):
    path, cfg, sd = _write_ckpt(tmp_path, "ab", "finetune")
    monkeypatch.setenv("HUDIFF_PACK", "0")
    _, state = S.load_checkpoint(path, "finetune", "ab")
    _same(state, sd)
    assert not os.path.exists(path + ".hudiffpack")
    monkeypatch.setenv("HUDIFF_PACK", "1")
    monkeypatch.setenv("HUDIFF_PACK_DIR", str(tmp_path / "does" / "not" / "exist"))
    _, state = S.load_checkpoint(path, "finetune", "ab")      # cannot write there: plain load, no error
    _same(state, sd)
    monkeypatch.setenv("HUDIFF_PACK_DIR", str(tmp_path))
    S.load_checkpoint(path, "finetune", "ab")
    assert os.path.exists(str(tmp_path / (os.path.basename(path) + ".hudiffpack")))
