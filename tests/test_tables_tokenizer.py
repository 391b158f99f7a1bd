"""Host pre/post-processing constants and tokenizer semantics."""
import ast
import os

import pytest
import torch

from hudiff_b200 import tables as tb
from hudiff_b200.tokenizer import Tokenizer

REF = "/root/reference"


def _ref_assignments(path, names):
    tree = ast.parse(open(path).read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) \
                and node.targets[0].id in names:
            out[node.targets[0].id] = ast.literal_eval(node.value)
    return out


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_tables_equal_reference():
    names = ["HEAVY_POSITIONS_dict", "LIGHT_POSITIONS_dict", "HEAVY_CDR_INDEX", "LIGHT_CDR_INDEX",
             "HEAVY_CDR_KABAT_NO_VERNIER", "LIGHT_CDR_KABAT_NO_VERNIER", "INPAINT_HEAVY_CDR_INDEX"]
    ref = _ref_assignments(os.path.join(REF, "dataset/preprocess.py"), names)
    ref.update(_ref_assignments(os.path.join(REF, "dataset/oas_pair_dataset_new.py"),
                                ["HEAVY_REGION_INDEX", "LIGHT_REGION_INDEX"]))
    assert len(ref) == 9
    for k, v in ref.items():
        assert getattr(tb, k) == v, k


def test_table_invariants():
    assert sum(1 for v in tb.HEAVY_CDR_INDEX if v == 0) == 93
    assert sum(1 for v in tb.LIGHT_CDR_INDEX if v == 0) == 92
    assert sum(1 for v in tb.HEAVY_CDR_KABAT_NO_VERNIER if v == 0) == 79
    assert sum(1 for v in tb.LIGHT_CDR_KABAT_NO_VERNIER if v == 0) == 78
    assert sum(1 for v in tb.INPAINT_HEAVY_CDR_INDEX if v == 0) == 87
    assert tb.HEAVY_POSITIONS_dict["111A"] == 111 and tb.HEAVY_POSITIONS_dict["112A"] == 134
    assert tb.HEAVY_POSITIONS_dict["128"] == 151 and tb.LIGHT_POSITIONS_dict["127"] == 138
    assert tb.LIGHT_POSITIONS_dict["112F"] == 117
    assert max(tb.HEAVY_REGION_INDEX) == 6 and tb.HEAVY_REGION_INDEX == sorted(tb.HEAVY_REGION_INDEX)


def test_tokenizer_vocab_and_roundtrip():
    t = Tokenizer()
    assert t.n_toks == 23 and t.idx_pad == 21 and t.idx_msk == 22 and t.tok2idx("X") == 20
    assert [t.tok2idx(c) for c in "ACDY"] == [0, 1, 2, 19]
    seqs = ["EVQLVESGGGLVQPGGSLRLSCAASGFTFS", "QVQLQESGGGLVQPGG"]
    mat = t.seq2idx_batch(seqs)
    assert mat.shape == (2, 30) and int(mat[1, -1]) == 21
    assert t.idx2seq_batch(mat) == seqs                      # pads dropped (tokenizer.py:101-111)
    assert t.idx2seq_pad(mat[1]).endswith("-" * 14)
    assert [t.chain_type_idx(c) for c in "HLK"] == [0, 1, 2]
    with pytest.raises(TypeError):
        t.chain_type_idx("Q")
    with pytest.raises(KeyError):
        t.seq2idx("B")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_mirror_modules_match_reference_init_and_rng():
    from oracle import ref_stubs
    from hudiff_b200 import AB_RELEASE, NB_RELEASE, AntiTFNetB200, NanoAntiTFNetB200
    AntiTFNet, NanoAntiTFNet = ref_stubs.load_models()
    small = dict(n_encoder_layers=1, dual_layers=1, cs_layers=1)
    state = torch.get_rng_state()
    for cfg, Ref, Mine in ((dict(AB_RELEASE, **small), AntiTFNet, AntiTFNetB200),
                           (dict(NB_RELEASE, **small), NanoAntiTFNet, NanoAntiTFNetB200)):
        torch.manual_seed(4)
        r = Ref(**cfg)
        s_r = torch.get_rng_state()
        torch.manual_seed(4)
        m = Mine(**cfg)
        s_m = torch.get_rng_state()
        assert torch.equal(s_r, s_m)
        sr, sm = r.state_dict(), m.state_dict()
        assert list(sr) == list(sm)
        for k in sr:
            assert sr[k].dtype == sm[k].dtype and torch.equal(sr[k], sm[k]), k
        m.load_state_dict(sr)                                 # strict load of a reference state dict
    torch.set_rng_state(state)
