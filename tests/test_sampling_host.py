"""Host pre/post-processing against the reference's own builders."""
import os
import sys

import numpy as np
import pytest
import torch

from hudiff_b200 import sampling as S
from hudiff_b200 import tables as tb

REF = "/root/reference"
H = "EVQLVESGGGLVQPGGSLRLSCAASGFTFSSYAMSWVRQAPGKGLEWVSAISGSGGSTYYADSVKGRFTISRDNSKNTLYLQMNSLRAEDTAVYYCAKDRLSITIRPRYYGLDVWGQGTTVTVSS"
L = "DIQMTQSPSSLSASVGDRVTITCRASQSISSYLNWYQQKPGKAPKLLIYAASSLQSGVPSRFSGSGSGTDFTLTISSLQPEDFATYYCQQSYSTPLTFGGGTKVEIK"


def fake_number(seq):
    return S.sequential_numbering("H" if seq == H else "K")(seq)


def test_builders_shapes_and_masks():
    tokens, region, chain, batch, loc, tok = S.batch_input_element(H, L, 3, 0, finetune=True, number=fake_number)
    assert tokens.shape == (3, 291) and region.shape == (3, 291)
    assert chain.tolist() == [0, 0, 0, 2, 2, 2] and batch.tolist() == [0, 0, 0, 1, 1, 1]
    assert (tokens[:, loc] == 22).all() and int((tokens == 22).sum()) == 3 * len(loc)
    frame = np.array(tb.HEAVY_CDR_KABAT_NO_VERNIER + tb.LIGHT_CDR_KABAT_NO_VERNIER) == 0
    assert frame[loc].all()
    # pretrain mask: every framework slot (185 = 93 + 92), gaps included
    _, _, _, _, loc_p, _ = S.batch_input_element(H, L, 1, 0, finetune=False, number=fake_number)
    assert len(loc_p) == 185
    _, _, _, _, loc_e, _ = S.batch_equal_input_element(H, L, 1, 0, number=fake_number)
    assert set(loc_e) <= set(loc_p)
    t_nb, r_nb, loc_nb, _ = S.nano_batch_input_element(H, False, 2, number=fake_number)
    assert t_nb.shape == (2, 152) and (t_nb[:, loc_nb] == 22).all()
    _, _, loc_in, _ = S.nano_batch_input_element(H, True, 2, number=fake_number)
    assert len(loc_in) <= 87


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_builders_equal_reference():
    from oracle import ref_stubs
    ref_stubs.install()
    from unittest.mock import MagicMock
    for name in ("dataset.abnativ_alignment", "dataset.abnativ_alignment.align_and_clean", "dataset.abnativ_alignment.mybio",
                 "anarci", "lmdb", "Bio", "Bio.Seq", "Bio.SeqRecord", "Bio.SeqIO", "seaborn", "matplotlib",
                 "matplotlib.pyplot", "bs4", "abnumber"):
        if name not in sys.modules or name == "abnumber":
            sys.modules[name] = MagicMock()
    sys.path.insert(0, os.path.join(REF, "antibody_scripts"))
    try:
        import sample as ref_sample
    except Exception as e:                                      # a deeper third-party import we did not stub
        pytest.skip(f"reference sample.py not importable here: {e}")
    ref_sample.get_pad_seq = lambda seq: fake_number(seq)
    for finetune in (True, False):
        want = ref_sample.batch_input_element(H, L, batch_size=3, pad_region=0, finetune=finetune)
        got = S.batch_input_element(H, L, 3, 0, finetune=finetune, number=fake_number)
        for w, g in zip(want[:4], got[:4]):
            assert torch.equal(torch.as_tensor(w), torch.as_tensor(g))
        assert np.array_equal(want[4], got[4])


def test_identity_selection():
    parent = torch.randint(0, 20, (4, 291))
    reps = parent.clone()
    reps[0, :50] = (reps[0, :50] + 1) % 20
    reps[1, :5] = (reps[1, :5] + 1) % 20
    reps[2, :100] = (reps[2, :100] + 1) % 20
    reps[3, :20] = (reps[3, :20] + 1) % 20
    assert S.select_most_similar(parent, reps) == 1
    assert float(S.slot_identity(parent[:1], parent[:1])) == 1.0


def test_fasta_readers_and_sample_until_semantics(tmp_path):
    """FASTA selection rules of the single-antibody scripts and the `while sample_number > 0` loop:
    every replica consumes one unit, duplicates are dropped but counted, tokens are not re-masked
    between passes (sample_for_anti_cdr.py:177-214, sample_for_nano_cdr.py:165-201)."""
    import torch
    from hudiff_b200 import sampling as S
    fa = tmp_path / "7k9i.fasta"
    fa.write_text(">7K9I_1|Chain A|Spike|virus\nNITNLC\n>7K9I_2|Chains B|Fab heavy chain|Homo\nEVQLVE\nSGGG\n"
                  ">7K9I_3|Chains C|Fab light chain|Homo\nDIQMTQ\n")
    assert S.h_l_from_fasta(str(fa)) == ("EVQLVESGGG", "DIQMTQ")
    nb = tmp_path / "nb.fasta"
    nb.write_text(">x|Chain A|Nanobody 7|Lama\nQVQLQE\n>x|Chain B|antigen\nAAA\n")
    assert S.nano_from_fasta(str(nb)) == "QVQLQE"

    class Fake:                                    # pass k returns rows [k, k, k+1]: duplicates inside and across passes
        def __init__(self):
            self.calls, self.inputs = 0, []

        def sample(self, tokens, region, chain, order, rng, seed):
            self.inputs.append(tokens.clone())
            k = self.calls
            self.calls += 1
            return torch.tensor([[k], [k], [k + 1]])

    m = Fake()
    start = torch.tensor([[22], [22], [22]])
    unique, passes = S.sample_until(m, start, start, None, [0], 7, lambda t: [int(r[0]) for r in t], device="cpu")
    assert passes == 3 and unique == [0, 1, 2, 3][:len(unique)] and unique == [0, 1, 2]   # 3 + 3 + 1 replicas consumed
    assert torch.equal(m.inputs[0], start) and torch.equal(m.inputs[1], torch.tensor([[0], [0], [1]]))  # fed back, not re-masked
    rejected = []

    def validate(res):
        if res == 1:
            rejected.append(res)
            raise ValueError("not an antibody")
    with pytest.raises(ValueError):
        S.sample_until(Fake(), start, start, None, [0], 3, lambda t: [int(r[0]) for r in t], device="cpu", validate=validate)
    assert rejected == [1]


# --------------------------------------------------------------------------------------------- #
# inpaint path, retry loops, best-replica selection
# --------------------------------------------------------------------------------------------- #
def fake_graft(seq):
    """Stand-in for abnumber's CDR grafting: 'germline' = parent with every 5th framework residue changed;
    kept keys = CDR positions + framework positions where parent == germline (what graft_chain returns)."""
    d, ctype = fake_number(seq)
    table = tb.HEAVY_POSITIONS_dict if ctype == "H" else tb.LIGHT_POSITIONS_dict
    cdr = tb.HEAVY_CDR_INDEX if ctype == "H" else tb.LIGHT_CDR_INDEX
    graft, keep = {}, []
    for i, (key, aa) in enumerate(d.items()):
        slot = table.get(key)
        in_cdr = slot is not None and cdr[slot] != 0
        changed = (not in_cdr) and i % 5 == 0
        graft[key] = "A" if (changed and aa != "A") else ("G" if changed else aa)
        if in_cdr or not changed:
            keep.append(key)
    return graft, keep, ctype


def _import_reference_sample():
    from oracle import ref_stubs
    ref_stubs.install()
    from unittest.mock import MagicMock
    for name in ("dataset.abnativ_alignment", "dataset.abnativ_alignment.align_and_clean", "dataset.abnativ_alignment.mybio",
                 "anarci", "lmdb", "Bio", "Bio.Seq", "Bio.SeqRecord", "Bio.SeqIO", "seaborn", "matplotlib",
                 "matplotlib.pyplot", "bs4", "abnumber"):
        if name not in sys.modules or name == "abnumber":
            sys.modules[name] = MagicMock()
    if os.path.join(REF, "antibody_scripts") not in sys.path:
        sys.path.insert(0, os.path.join(REF, "antibody_scripts"))
    try:
        import sample as ref_sample
    except Exception as e:                                      # a deeper third-party import we did not stub
        pytest.skip(f"reference sample.py not importable here: {e}")
    return ref_sample


def test_inpaint_builder_masks():
    tokens, region, chain, batch, loc, tok = S.batch_inpaint_input_element(H, L, 2, 0, graft=fake_graft)
    assert tokens.shape == (2, 291) and chain.tolist() == [0, 0, 2, 2]
    frame = np.array(tb.HEAVY_CDR_INDEX + tb.LIGHT_CDR_INDEX) == 0
    assert frame[loc].all() and (tokens[:, loc] == 22).all() and int((tokens == 22).sum()) == 2 * len(loc)
    # CDR residues of the graft are kept, the changed framework residues are sampled
    parent = tok.seq2idx(S.get_input_element(H, L, 0, fake_number)[1])
    cdr = ~frame
    assert torch.equal(tokens[0][torch.from_numpy(cdr)], parent[torch.from_numpy(cdr)])
    assert len(loc) >= (len(H) + len(L)) // 5 - 10


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_inpaint_builder_equals_reference():
    ref_sample = _import_reference_sample()
    ref_sample.graft_chain = fake_graft
    for bs in (1, 3):
        want = ref_sample.batch_inpaint_input_element(H, L, batch_size=bs, pad_region=0)
        got = S.batch_inpaint_input_element(H, L, bs, 0, graft=fake_graft)
        for w, g in zip(want[:4], got[:4]):
            assert torch.equal(torch.as_tensor(w), torch.as_tensor(g))
        assert np.array_equal(want[4], got[4])


class FakeAlign:
    def __init__(self, d1, d2):
        self.d1, self.d2 = d1, d2
        self.positions = sorted(set(d1) | set(d2))

    def __getitem__(self, pos):
        return self.d1.get(pos, "-"), self.d2.get(pos, "-")


class FakeChain:
    """abnumber.Chain stand-in: gap-free numbering, align() over the union of positions."""

    def __init__(self, seq, scheme="imgt"):
        self.seq = seq
        self.d = {i + 1: a for i, a in enumerate(seq)}

    def align(self, other):
        return FakeAlign(self.d, other.d)


def _mutate(seq, idx, to="W"):
    s = list(seq)
    for i in idx:
        s[i] = to if s[i] != to else "Y"
    return "".join(s)


def test_best_replica_selection_matches_reference_function():
    hs = [_mutate(H, range(0, 40)), _mutate(H, range(0, 6)), _mutate(H, range(0, 6)), H[:-3]]
    ls = [_mutate(L, range(0, 10)), _mutate(L, range(0, 30)), _mutate(L, range(0, 4)), L]
    k, g_h, g_l = S.select_the_most_similarity_seq(H, L, hs, ls, chain_factory=FakeChain)
    assert k == 3 and (g_h, g_l) == (hs[3], ls[3])           # three missing residues beat ten substitutions
    # slot-grid identity picks the same replica when every residue sits on the grid
    tok = S.Tokenizer()
    grid = lambda h, l: tok.seq2idx(S.get_input_element(h, l, 0, lambda s: S.sequential_numbering("H")(s))[1])   # noqa: E731
    parent = grid(H, L).unsqueeze(0).expand(4, -1)
    reps = torch.stack([grid(h, l) for h, l in zip(hs, ls)])
    assert S.select_most_similar(parent, reps) == k
    if os.path.isdir(REF):
        ref_sample = _import_reference_sample()
        ref_sample.Chain = FakeChain
        assert ref_sample.select_the_most_similarity_seq(H, L, hs, ls) == (g_h, g_l)


class _LoopModel:
    """model.sample stand-in: pass k turns token row b into the value 10 * k + b."""

    def __init__(self):
        self.calls, self.inputs, self.offsets = 0, [], []

    def sample(self, tokens, region, chain, order, rng, seed, seq_offset=0):
        self.inputs.append(tokens.clone())
        self.offsets.append(seq_offset)
        out = tokens.clone()
        out[:, 0] = torch.arange(tokens.shape[0]) % 20
        out[:, 1] = self.calls % 20
        self.calls += 1
        return out


def test_ab_pass_loop_semantics():
    """sample.py:499-536: no re-masking between passes; similarity search writes one row and stops; otherwise every
    replica consumes one unit of sample_number; try_num is not decremented by the Ab script."""
    tok = S.Tokenizer()
    start = torch.full((3, 291), 21)
    start[:, :2] = 22
    region = torch.zeros((3, 291), dtype=torch.long)
    chain = torch.tensor([0, 0, 0, 2, 2, 2])
    m = _LoopModel()
    rows, out, passes = S.ab_sample_passes(m, start, region, chain, [0, 1], 7, 1, False, device="cpu")
    assert passes == 3 and len(rows) == 7                        # 3 + 3 + 1 replicas written
    assert torch.equal(m.inputs[0], start) and (m.inputs[1][:, :2] != 22).all()      # completed tokens are fed back
    assert m.offsets == [0, 3, 6]                                # fresh philox streams per pass
    assert rows[0] == (tok.idx2seq(torch.tensor([0, 0])), "") and rows[3] == (tok.idx2seq(torch.tensor([0, 1])), "")
    m = _LoopModel()
    rows, _, passes = S.ab_sample_passes(m, start, region, chain, [0, 1], 5, 1, True, select=lambda t, h, l: 2, device="cpu")
    assert passes == 1 and rows == [(tok.idx2seq(torch.tensor([2, 0])), "")]
    rows, _, passes = S.ab_sample_passes(_LoopModel(), start, region, chain, [0, 1], 2, 0, False, device="cpu")
    assert rows == [] and passes == 0                            # try_number = 0 never enters the loop


def test_nano_pass_loop_semantics():
    """nanosample.py:316-353: try_num counts every replica looked at; an invalid replica is written only when
    try_num == 1; a valid one consumes sample_number."""
    start = torch.full((2, 152), 21)
    start[:, :2] = 22
    region = torch.zeros((2, 152), dtype=torch.long)
    seen = []

    def validate(seq):                                           # replica 1 of every pass is "not an antibody"
        seen.append(seq)
        if seq[0] == "C":
            raise ValueError("cannot be numbered")
    rows, passes = S.nano_sample_passes(_LoopModel(), start, region, [0, 1], 3, 10, validate, device="cpu")
    assert passes == 3 and rows == ["AA", "AC", "AD"] and len(seen) == 5      # A? valid, C? rejected, 3rd pass stops early
    rows, passes = S.nano_sample_passes(_LoopModel(), start, region, [0, 1], 3, 2, validate, device="cpu")
    assert passes == 1 and rows == ["AA", "CA"]                  # the invalid replica is written on the last try
    rows, passes = S.nano_sample_passes(_LoopModel(), start, region, [0, 1], 5, 3, None, device="cpu")
    assert passes == 2 and len(rows) == 4                        # try_num is only tested between passes (:316), like the reference


def test_number_batch_numbers_each_sequence_once_and_keeps_errors():
    calls = []

    def slow_number(seq):
        calls.append(seq)
        if seq == "BAD":
            raise ValueError("not an antibody")
        return S.sequential_numbering("H")(seq)
    fn = S.number_batch([H, L, H, "BAD", L], slow_number, workers=4)
    assert sorted(calls) == sorted([H, L, "BAD"])                # every distinct sequence exactly once
    assert fn(H) == S.sequential_numbering("H")(H) and fn(L)[0]["1"] == L[0]
    d, _ = fn(H)
    d["1"] = "?"                                                 # callers get copies
    assert fn(H)[0]["1"] == H[0]
    with pytest.raises(ValueError):
        fn("BAD")
    assert len(calls) == 3
    tokens, *_ = S.batch_input_element(H, L, 2, 0, True, number=lambda s: (fn(s)[0], "H" if s == H else "K"))
    assert tokens.shape == (2, 291)
