"""BASELINE config 1 (reference plumbing): the drop-in `antibody_scripts/sample.py` and
`nanobody_scripts/nanosample.py` on 4 HuAb348 mouse antibodies, checkpoints in the reference's formats.

ANARCI / abnumber are unavailable offline, so numbering comes from a fixture (gap-free stand-in); parity
is pinned at the token-tensor boundary: the CLI output must equal the CPU oracle driven through the same
host pipeline under the same seed (torch_cpu_replay, live eval-mode dropout)."""
import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from hudiff_b200 import AB_RELEASE, NB_RELEASE, EasyDict, seed_all
from hudiff_b200 import sampling as S
from oracle import hudiff_oracle as O
from util import make_model

def _load_cli(subdir, name):
    """Import a drop-in script by path (the reference's own `sample` module may already sit in sys.modules)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(f"hudiff_cli_{name}", os.path.join(ROOT, subdir, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


SMALL = dict(n_encoder_layers=1, dual_layers=1, cs_layers=1)
CSV = os.path.join(ROOT, "tests", "golden", "huab348_first4_mouse.csv")
NUMBERING = os.path.join(ROOT, "tests", "golden", "huab348_first4_numbering.json")


def _write_ckpt(tmp_path, kind, version):
    model, cfg = make_model(kind, dict(AB_RELEASE if kind == "ab" else NB_RELEASE, **SMALL), seed=3)
    sd = model.state_dict()
    config = EasyDict(name="trans_oadm" if kind == "ab" else "nano", model=cfg)
    if kind == "ab" and version == "finetune":            # antibody_finetune.py:348-355
        ckpt = {"fineconfig": EasyDict(), "pretrain_config": config, "model": sd, "optimizer": {}, "scheduler": {}, "iteration": 1}
    elif kind == "nb" and version == "finetune_vh":       # nanofinetune.py:531-539
        full = {"infilling_pretrain." + k: v for k, v in sd.items()}
        full.update({"target_infilling_pretrain." + k: v for k, v in sd.items()})
        full["eval_abnativ_model.dummy"] = torch.zeros(1)
        ckpt = {"config": config, "model": full, "abnativ_params": {}, "infilling_params": cfg}
    else:                                                 # antibody_train.py:439-445 (+ DataParallel prefix)
        ckpt = {"config": config, "model": {"module." + k: v for k, v in sd.items()}, "optimizer": {}, "scheduler": {}, "iteration": 1}
    d = tmp_path / "ckpt" / "run"
    d.mkdir(parents=True)
    path = str(d / f"{kind}_{version}.pt")
    torch.save(ckpt, path)
    return path, cfg, sd


@pytest.mark.parametrize("kind,version", [("ab", "finetune"), ("ab", "pretrain"), ("nb", "finetune_vh"), ("nb", "pretrain")])
def test_checkpoint_formats_load(tmp_path, kind, version):
    path, cfg, sd = _write_ckpt(tmp_path, kind, version)
    config, state = S.load_checkpoint(path, version, kind)
    model_cfg = config["model"] if isinstance(config, dict) else config.model
    assert dict(model_cfg) == cfg
    assert set(state) == set(sd)
    for k in sd:
        assert torch.equal(state[k], sd[k])


@pytest.mark.gpu
def test_sample_cli_equals_oracle_pipeline(tmp_path):
    cli = _load_cli("antibody_scripts", "sample")
    path, cfg, sd = _write_ckpt(tmp_path, "ab", "finetune")
    out_dir = str(tmp_path / "out")
    seed, B = 0, 2
    save = cli.main(["--ckpt", path, "--ckpt_version", "finetune", "--data_fpath", CSV, "--batch_size", str(B),
                     "--seed", str(seed), "--numbering_json", NUMBERING, "--rng", "torch_cpu_replay",
                     "--precision", "fp16x3", "--out_dir", out_dir])
    got = pd.read_csv(save, index_col=False)
    assert list(got.columns)[:4] == ["Specific", "name", "hseq", "lseq"]
    human = got[got["Specific"] == "humanization"].reset_index(drop=True)
    mouse = got[got["Specific"] == "mouse"].reset_index(drop=True)
    assert len(human) == 4 and len(mouse) == 4
    assert open(os.path.join(out_dir, "sample_identity.fa")).read().count(">") == 8
    # the same pipeline on the CPU oracle: same seed discipline as the CLI (seed_all, model construction, loop)
    seed_all(seed)
    model, _ = make_model("ab", cfg, seed=3)       # consumes no global RNG (state restored); the CLI's model ctor does
    from hudiff_b200 import model_selected
    model_selected(EasyDict(name="trans_oadm", model=cfg))            # replicate the CLI's RNG consumption
    number = S.FixtureNumbering.from_json(NUMBERING)
    df = pd.read_csv(CSV)
    for idx, line in enumerate(df.itertuples()):
        tokens, region, chain, _, loc, tok = S.batch_input_element(line.h_seq, line.l_seq, B, 0, True, number)
        parent = tok.seq2idx(S.get_input_element(line.h_seq, line.l_seq, 0, number)[1]).unsqueeze(0).expand(B, -1)
        np.random.shuffle(loc)
        out = O.sample_loop("ab", sd, cfg, tokens, region, chain, torch.from_numpy(loc.copy()), O.TorchStream())
        h, l = S.split_and_untokenize(out, tok)
        k = S.select_most_similar(parent, out)
        assert human.loc[idx, "hseq"] == h[k] and human.loc[idx, "lseq"] == l[k], f"antibody {idx}"


@pytest.mark.gpu
def test_nanosample_cli_runs(tmp_path):
    cli = _load_cli("nanobody_scripts", "nanosample")
    path, cfg, sd = _write_ckpt(tmp_path, "nb", "finetune_vh")
    df = pd.read_csv(CSV)
    vhh_csv = str(tmp_path / "vhh.csv")
    pd.DataFrame({"vhhseq": df["h_seq"]}).to_csv(vhh_csv, index=False)
    save = cli.main(["--ckpt", path, "--model", "finetune_vh", "--data_fpath", vhh_csv, "--batch_size", "2",
                     "--numbering_json", NUMBERING, "--inpaint_sample", "True", "--out_dir", str(tmp_path / "o")])
    got = pd.read_csv(save, index_col=False)
    human = got[got["Specific"] == "humanization"]
    assert len(human) == 4 and all(set(s) <= set("ACDEFGHIKLMNPQRSTVWYX") for s in human["hseq"])


@pytest.mark.gpu
def test_single_antibody_front_ends(tmp_path):
    """`sample_for_anti_cdr.py` / `sample_for_nano_cdr.py`: CSV layout, de-duplication,
    and -- antibody script, torch_cpu_replay -- equality with the CPU oracle driven through the same two passes."""
    df = pd.read_csv(CSV)
    h_seq, l_seq = df.loc[0, "h_seq"], df.loc[0, "l_seq"]
    fa = tmp_path / "1abc.fasta"
    fa.write_text(f">1ABC_1|Chain H|Fab heavy chain|Mus\n{h_seq}\n>1ABC_2|Chain L|Fab light chain|Mus\n{l_seq}\n")
    path, cfg, sd = _write_ckpt(tmp_path, "ab", "finetune")
    cli = _load_cli("antibody_scripts", "sample_for_anti_cdr")
    seed, B, N = 1, 2, 4                                  # two passes of two replicas
    save = cli.main(["--ckpt", path, "--anti_complex_fasta", str(fa), "--batch_size", str(B), "--sample_number", str(N),
                     "--seed", str(seed), "--numbering_json", NUMBERING, "--rng", "torch_cpu_replay",
                     "--precision", "fp16x3", "--out_dir", str(tmp_path / "ab_out")])
    got = pd.read_csv(save, index_col=False)
    assert list(got.columns)[:4] == ["Specific", "name", "hseq", "lseq"]
    assert got.loc[0, "Specific"] == "mouse" and got.loc[0, "name"] == "1abc" and got.loc[0, "hseq"] == h_seq
    human = got[got["Specific"] == "humanization"].reset_index(drop=True)
    assert 1 <= len(human) <= N and not human.duplicated(["hseq", "lseq"]).any()
    assert (human["name"] == "1abchuman_sample").all()
    # oracle: same seed discipline (seed_all, model construction, shuffle, two passes without re-masking)
    seed_all(seed)
    from hudiff_b200 import model_selected
    model_selected(EasyDict(name="trans_oadm", model=cfg))
    number = S.FixtureNumbering.from_json(NUMBERING)
    tokens, region, chain, _, loc, tok = S.batch_input_element(h_seq, l_seq, B, 0, True, number)
    np.random.shuffle(loc)
    want, seen = [], set()
    for _ in range(N // B):
        tokens = O.sample_loop("ab", sd, cfg, tokens, region, chain, torch.from_numpy(loc.copy()), O.TorchStream())
        for pair in zip(*S.split_and_untokenize(tokens, tok)):
            if pair not in seen:
                seen.add(pair)
                want.append(pair)
    assert [tuple(r) for r in human[["hseq", "lseq"]].itertuples(index=False)] == want

    nb_fa = tmp_path / "9nb.fasta"
    nb_fa.write_text(f">9NB_1|Chain A|Nanobody Nb1|Lama glama\n{h_seq}\n>9NB_2|Chain B|antigen\nAAAA\n")
    npath, ncfg, nsd = _write_ckpt(tmp_path / "n", "nb", "finetune_vh")
    ncli = _load_cli("nanobody_scripts", "sample_for_nano_cdr")
    nsave = ncli.main(["--ckpt", npath, "--nano_complex_fasta", str(nb_fa), "--batch_size", "3", "--sample_number", "5",
                       "--numbering_json", NUMBERING, "--out_dir", str(tmp_path / "nb_out")])
    ngot = pd.read_csv(nsave, index_col=False)
    assert list(ngot.columns)[:3] == ["Specific", "name", "hseq"] and ngot.loc[0, "Specific"] == "Nano"
    nh = ngot[ngot["Specific"] == "humanization"]
    assert 1 <= len(nh) <= 5 and not nh["hseq"].duplicated().any()
    assert open(os.path.join(str(tmp_path / "nb_out"), "sample_identity.fa")).read().count(">") == len(nh)
