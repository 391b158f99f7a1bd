"""Amino-acid tokenizer with the reference's vocabulary (utils/tokenizer.py:43-149).

23 tokens: the 20 residues in alphabetical one-letter order (0..19), 'X' (20),
pad '-' (21) and '<msk>' (22).  `idx2seq` drops pad tokens, `idx2seq_pad` keeps
them; `chain_type_idx` maps H/L/K to 0/1/2 and raises TypeError otherwise.
"""
from __future__ import annotations

import torch

AA20 = "ACDEFGHIKLMNPQRSTVWY"


class Tokenizer:
    tok_pad = "-"
    tok_msk = "<msk>"

    def __init__(self, has_bos: bool = False, has_eos: bool = False):
        if has_bos or has_eos:
            raise NotImplementedError("the sampling path never uses bos/eos tokens")
        self.has_bos = has_bos
        self.has_eos = has_eos
        self.toks = list(AA20) + ["X", self.tok_pad, self.tok_msk]
        self.tok2idx_dict = {t: i for i, t in enumerate(self.toks)}
        self.idx_pad = self.tok2idx_dict[self.tok_pad]
        self.idx_msk = self.tok2idx_dict[self.tok_msk]

    @property
    def n_toks(self) -> int:
        return len(self.toks)

    def tok2idx(self, tok: str) -> int:
        return self.tok2idx_dict[tok]

    def seq2idx(self, aa_seq) -> torch.Tensor:
        return torch.tensor([self.tok2idx_dict[c] for c in aa_seq])

    def seq2idx_batch(self, seqs) -> torch.Tensor:
        rows = [self.seq2idx(s) for s in seqs]
        return torch.nn.utils.rnn.pad_sequence(rows, batch_first=True, padding_value=self.idx_pad)

    def idx2seq(self, idx_vec) -> str:
        return "".join(self.toks[i] for i in torch.as_tensor(idx_vec).tolist() if i != self.idx_pad)

    def idx2seq_pad(self, idx_vec) -> str:
        return "".join(self.toks[i] for i in torch.as_tensor(idx_vec).tolist())

    def idx2seq_batch(self, idx_mat):
        return [self.idx2seq(r) for r in idx_mat]

    def idx2seq_pad_batch(self, idx_mat):
        return [self.idx2seq_pad(r) for r in idx_mat]

    def chain_type_idx(self, chain: str) -> int:
        try:
            return {"H": 0, "L": 1, "K": 2}[chain]
        except KeyError:
            raise TypeError("Chain Type has problem.") from None
