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
