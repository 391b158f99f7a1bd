"""Host pre/post-processing of the sampling path and the humanization drivers.

Mirrors the reference's input builders (antibody_scripts/sample.py:94-206,
nanobody_scripts/nanosample.py:91-149): IMGT-numbered residues are scattered onto the fixed slot
grid (152 heavy + 139 light), framework slots are masked (token 22) and the masked slot indices
become the sampling order.  The denoising loop itself runs on the GPU (`model.sample`).

Numbering is third-party in the reference (ANARCI + abnumber, unavailable offline): `number_sequence`
uses them when importable; otherwise a `FixtureNumbering` (JSON of pre-numbered sequences) must be
supplied.  Everything downstream of the numbering dict is implemented here.
"""
from __future__ import annotations

import json
import re
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import tables as tb
from .tokenizer import Tokenizer

NumberFn = Callable[[str], Tuple[Dict[str, str], str]]      # seq -> ({'<imgt pos><ins>': aa}, chain type H/L/K)


# --------------------------------------------------------------------------------------------- #
# numbering back-ends
# --------------------------------------------------------------------------------------------- #
def anarci_numbering(aa_seq: str):
    """sample.py:78-90 get_pad_seq -- ANARCI IMGT numbering + abnumber chain type."""
    try:
        from abnumber import Chain
        from anarci import number
    except ImportError as e:                                      # pragma: no cover - not installed offline
        raise RuntimeError(
            "ANARCI / abnumber are not installed; pass a FixtureNumbering (see --numbering_json)") from e
    results = number(aa_seq, scheme="imgt")
    seq_dict = {str(key[0]) + key[1].strip(): value for key, value in results[0]}
    return seq_dict, Chain(aa_seq, scheme="imgt").chain_type


class FixtureNumbering:
    """Pre-computed numbering: {sequence: {"positions": {"1": "E", "111A": "G", ...}, "chain_type": "H"}}."""

    def __init__(self, table: Dict[str, dict]):
        self.table = table

    @classmethod
    def from_json(cls, path: str):
        with open(path) as f:
            return cls(json.load(f))

    def __call__(self, aa_seq: str):
        try:
            rec = self.table[aa_seq]
        except KeyError:
            raise KeyError(f"sequence not in the numbering fixture: {aa_seq[:20]}...") from None
        return dict(rec["positions"]), rec["chain_type"]


def number_batch(seqs, number: NumberFn = anarci_numbering, workers: int = 0) -> NumberFn:
    """Number many sequences up front (at thousands of sequences per second on the GPU the
    per-chain ANARCI call -- an hmmscan subprocess, 0.1-1 s -- dominates end to end) and return a memoised
    NumberFn for the builders.  Every distinct sequence is numbered exactly once, `workers` threads overlap
    the subprocess waits (0 -> one per host core, capped by the number of distinct sequences); errors are
    re-raised on lookup, in the caller's order."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    distinct = list(dict.fromkeys(seqs))
    n = workers or (os.cpu_count() or 1)
    results: Dict[str, object] = {}

    def one(seq):
        try:
            return seq, number(seq)
        except Exception as e:                                   # surfaced when the sequence is looked up
            return seq, e
    if distinct:
        with ThreadPoolExecutor(max_workers=max(1, min(n, len(distinct)))) as pool:
            for seq, res in pool.map(one, distinct):
                results[seq] = res

    def lookup(aa_seq: str):
        res = results[aa_seq] if aa_seq in results else one(aa_seq)[1]
        if isinstance(res, Exception):
            raise res
        d, ctype = res
        return dict(d), ctype
    return lookup


def sequential_numbering(chain_type: str) -> NumberFn:
    """Gap-free stand-in (residue i -> IMGT position i+1): plumbing tests only, NOT a real numbering."""
    def fn(aa_seq: str):
        return {str(i + 1): a for i, a in enumerate(aa_seq)}, chain_type
    return fn


# --------------------------------------------------------------------------------------------- #
# slot scatter (sample.py:94-139, nanosample.py:91-121)
# --------------------------------------------------------------------------------------------- #
_CDR_RANGES = ((27, 38), (56, 65), (105, 117))


def _scatter(seq_dict: Dict[str, str], pos_table: Dict[str, int], n_slots: int, label: str, log=print,
             keep=None) -> List[str]:
    """Residues onto the slot grid.  `keep` (inpaint path, sample.py:250-254): only these position keys are placed."""
    slots = ["-"] * n_slots
    for key, value in seq_dict.items():
        if keep is not None and key not in keep:
            continue
        idx = pos_table.get(key)
        if idx is not None:
            slots[idx] = value
            continue
        nkey = int(re.findall(r"\d+", key)[0])
        if any(lo <= nkey <= hi for lo, hi in _CDR_RANGES):
            log(f"{label} CDR has problem.")
        else:
            log(f"{label[0]} Position {key} is not in predefine dict, which can be ignored.")
    return slots


def get_input_element(mouse_aa_h: str, mouse_aa_l: str, pad_region: int = 0, number: NumberFn = anarci_numbering):
    h_dict, h_type = number(mouse_aa_h)
    l_dict, l_type = number(mouse_aa_l)
    region = torch.cat((torch.tensor(tb.HEAVY_REGION_INDEX), torch.tensor(tb.LIGHT_REGION_INDEX) + pad_region))
    slots = _scatter(h_dict, tb.HEAVY_POSITIONS_dict, tb.H_LEN, "Heavy") + \
        _scatter(l_dict, tb.LIGHT_POSITIONS_dict, tb.L_LEN, "Light")
    return region, slots, [h_type, l_type], torch.tensor([0, 1])


def _expand(tokens, region, chain_type, ms_batch, batch_size, tok):
    region = region.unsqueeze(0).expand(batch_size, -1).clone()
    tokens = tokens.unsqueeze(0).expand(batch_size, -1).clone()
    chain = torch.tensor([tok.chain_type_idx(c) for c in chain_type]).view(-1, 1).repeat(1, batch_size).view(-1)
    ms_batch = ms_batch.view(-1, 1).repeat(1, batch_size).view(-1)
    return tokens, region, chain, ms_batch


def batch_input_element(mouse_sq_h, mouse_sq_l, batch_size=10, pad_region=0, finetune=False,
                        number: NumberFn = anarci_numbering):
    """sample.py:142-179 -- FR sampling inputs: B replicas of one antibody + masked slot indices."""
    region, slots, chain_type, ms_batch = get_input_element(mouse_sq_h, mouse_sq_l, pad_region, number)
    table = (tb.HEAVY_CDR_KABAT_NO_VERNIER + tb.LIGHT_CDR_KABAT_NO_VERNIER) if finetune \
        else (tb.HEAVY_CDR_INDEX + tb.LIGHT_CDR_INDEX)
    mask = torch.tensor(table) == 0
    loc = np.arange(tb.AB_LEN)
    tok = Tokenizer()
    tokens = tok.seq2idx(slots)
    if finetune:                                               # framework gaps are not sampled (sample.py:163-167)
        mask = mask & ~(tokens == tok.idx_pad)
    tokens[mask] = tok.idx_msk
    loc = loc[mask.numpy()]
    tokens, region, chain, ms_batch = _expand(tokens, region, chain_type, ms_batch, batch_size, tok)
    return tokens, region, chain, ms_batch, loc, tok


def batch_equal_input_element(mouse_sq_h, mouse_sq_l, batch_size=10, pad_region=0, number: NumberFn = anarci_numbering):
    """sample.py:182-206 -- pretrain mask restricted to non-gap slots."""
    region, slots, chain_type, ms_batch = get_input_element(mouse_sq_h, mouse_sq_l, pad_region, number)
    tok = Tokenizer()
    tokens = tok.seq2idx(slots)
    mask = (torch.tensor(tb.HEAVY_CDR_INDEX + tb.LIGHT_CDR_INDEX) == 0) & (tokens != tok.idx_pad)
    loc = np.arange(tb.AB_LEN)[mask.numpy()]
    tokens[mask] = tok.idx_msk
    tokens, region, chain, ms_batch = _expand(tokens, region, chain_type, ms_batch, batch_size, tok)
    return tokens, region, chain, ms_batch, loc, tok


# --------------------------------------------------------------------------------------------- #
# inpaint path: CDRs grafted onto the closest human germline, non-identical framework slots sampled
# (sample.py:209-310, --sample_method inpaint)
# --------------------------------------------------------------------------------------------- #
GraftFn = Callable[[str], Tuple[Dict[str, str], List[str], str]]   # seq -> (graft numbering dict, identity keys, chain type)


def abnumber_graft_chain(aa_seq: str, number: NumberFn = anarci_numbering):
    """sample.py:209-227 graft_chain -- abnumber CDR grafting; returns the numbering dict of the grafted chain,
    the position keys kept from it (every CDR position + framework positions where parent == germline) and
    the chain type."""
    try:
        from abnumber import Chain
    except ImportError as e:                                      # pragma: no cover - not installed offline
        raise RuntimeError("abnumber is not installed; pass a FixtureGraft (see --graft_json)") from e
    seq_chain = Chain(aa_seq, scheme="imgt")
    graft = seq_chain.graft_cdrs_onto_human_germline()
    align = seq_chain.align(graft)
    identity = []
    for pos in align.positions:
        if not pos.is_in_cdr():
            a1, a2 = align[pos]
            if a1 == a2:
                identity.append(str(pos)[1:])
        else:
            identity.append(str(pos)[1:])
    seq_dict, chain_type = number(graft.seq)
    return seq_dict, identity, chain_type


class FixtureGraft:
    """Pre-computed grafts: {sequence: {"positions": {...}, "identity": [...], "chain_type": "H"}}."""

    def __init__(self, table: Dict[str, dict]):
        self.table = table

    @classmethod
    def from_json(cls, path: str):
        with open(path) as f:
            return cls(json.load(f))

    def __call__(self, aa_seq: str):
        try:
            rec = self.table[aa_seq]
        except KeyError:
            raise KeyError(f"sequence not in the graft fixture: {aa_seq[:20]}...") from None
        return dict(rec["positions"]), list(rec["identity"]), rec["chain_type"]


def get_inpaint_input(mouse_aa_h: str, mouse_aa_l: str, pad_region: int = 0, graft: GraftFn = abnumber_graft_chain, log=print):
    """sample.py:230-283 -- slot grid holding only the grafted residues that are kept (CDRs + germline-identical
    framework); every other slot stays '-'."""
    h_dict, h_keep, h_type = graft(mouse_aa_h)
    l_dict, l_keep, l_type = graft(mouse_aa_l)
    region = torch.cat((torch.tensor(tb.HEAVY_REGION_INDEX), torch.tensor(tb.LIGHT_REGION_INDEX) + pad_region))
    slots = _scatter(h_dict, tb.HEAVY_POSITIONS_dict, tb.H_LEN, "Heavy", log, keep=h_keep) + \
        _scatter(l_dict, tb.LIGHT_POSITIONS_dict, tb.L_LEN, "Light", log, keep=l_keep)
    return region, slots, [h_type, l_type], torch.tensor([0, 1])


def batch_inpaint_input_element(mouse_sq_h, mouse_sq_l, batch_size=1, pad_region=0, graft: GraftFn = abnumber_graft_chain):
    """sample.py:286-310 -- sampled slots = framework slots (pretrain table) left empty by the graft."""
    region, slots, chain_type, ms_batch = get_inpaint_input(mouse_sq_h, mouse_sq_l, pad_region, graft)
    tok = Tokenizer()
    tokens = tok.seq2idx(slots)
    mask = (torch.tensor(tb.HEAVY_CDR_INDEX + tb.LIGHT_CDR_INDEX) == 0) & (tokens == tok.idx_pad)
    loc = np.arange(tb.AB_LEN)[mask.numpy()]
    tokens[mask] = tok.idx_msk
    tokens, region, chain, ms_batch = _expand(tokens, region, chain_type, ms_batch, batch_size, tok)
    return tokens, region, chain, ms_batch, loc, tok


def nano_batch_input_element(nano_sq, inpaint_sample=False, batch_size=10, number: NumberFn = anarci_numbering):
    """nanosample.py:124-149 -- VHH inputs (mask = framework slots that hold a residue)."""
    seq_dict = number(nano_sq)
    if isinstance(seq_dict, tuple):
        seq_dict = seq_dict[0]
    slots = _scatter(seq_dict, tb.HEAVY_POSITIONS_dict, tb.H_LEN, "Heavy")
    index = tb.INPAINT_HEAVY_CDR_INDEX if inpaint_sample else tb.HEAVY_CDR_INDEX
    tok = Tokenizer()
    tokens = tok.seq2idx(slots)
    frame = torch.tensor(index) == 0
    mask = frame & (tokens != tok.idx_pad)
    tokens[mask] = tok.idx_msk
    region = torch.tensor(tb.HEAVY_REGION_INDEX).unsqueeze(0).expand(batch_size, -1).clone()
    tokens = tokens.unsqueeze(0).expand(batch_size, -1).clone()
    loc = np.arange(tb.H_LEN)[mask.numpy()]
    return tokens, region, loc, tok


# --------------------------------------------------------------------------------------------- #
# drivers: one antibody / nanobody -> sampled replicas
# --------------------------------------------------------------------------------------------- #
def split_and_untokenize(tokens: torch.Tensor, tok: Tokenizer):
    """sample.py:515-518."""
    h = [tok.idx2seq(s) for s in tokens[:, :tb.H_LEN]]
    l = [tok.idx2seq(s) for s in tokens[:, tb.H_LEN:]]
    return h, l


def slot_identity(ref_tokens: torch.Tensor, tokens: torch.Tensor, pad: int = 21) -> torch.Tensor:
    """Mean per-slot identity to the parental sequence over slots where either holds a residue.

    On the fixed IMGT grid this equals the aligned-position identity of `cal_all_preservation`
    (antibody_scripts/patent_eval.py:150-159) without the per-replica abnumber alignment."""
    both = (ref_tokens != pad) | (tokens != pad)
    same = (ref_tokens == tokens) & both
    return same.sum(-1).float() / both.sum(-1).clamp(min=1).float()


def select_most_similar(ref_tokens: torch.Tensor, tokens: torch.Tensor) -> int:
    """Replica with the highest mean (heavy, light) identity to the parent (sample.py:352-367)."""
    hi = slot_identity(ref_tokens[:, :tb.H_LEN], tokens[:, :tb.H_LEN])
    li = slot_identity(ref_tokens[:, tb.H_LEN:], tokens[:, tb.H_LEN:])
    return int(torch.argmax((hi + li) / 2))


def chain_identity(chain1, chain2) -> float:
    """antibody_scripts/patent_eval.py:150-159 cal_all_preservation: identity over the aligned position union."""
    align = chain1.align(chain2)
    same = total = 0
    for pos in align.positions:
        a1, a2 = align[pos]
        same += int(a1 == a2)
        total += 1
    return same / total


def select_the_most_similarity_seq(ref_h_seq, ref_l_seq, h_untokenized, l_untokenized, chain_factory=None):
    """sample.py:352-367 -- replica with the highest mean (heavy, light) aligned identity to the parent.
    `chain_factory(seq) -> abnumber.Chain-like` (needs `.align`); default: abnumber's Chain(seq, scheme='imgt').
    Returns (index, h_seq, l_seq); ties resolve to the first replica like `list.index(max(...))`."""
    if chain_factory is None:
        try:
            from abnumber import Chain
        except ImportError as e:                                  # pragma: no cover - not installed offline
            raise RuntimeError("abnumber is not installed; use select_most_similar (slot-grid identity)") from e
        chain_factory = lambda seq: Chain(seq, scheme="imgt")     # noqa: E731
    ref_h, ref_l = chain_factory(ref_h_seq), chain_factory(ref_l_seq)
    ident = [(chain_identity(ref_h, chain_factory(h)) + chain_identity(ref_l, chain_factory(l))) / 2
             for h, l in zip(h_untokenized, l_untokenized)]
    k = ident.index(max(ident))
    return k, h_untokenized[k], l_untokenized[k]


def ab_sample_passes(model, tokens, region, chain, loc, sample_number: int, try_number: int, similarity_search: bool,
                     select=None, rng="philox", seed=0, device="cuda"):
    """The `while sample_number > 0 and try_num > 0` loop of sample.py:499-536 around the device loop.

    Faithful quirks: the tokens are NOT re-masked between passes (a later pass re-samples every slot of the already
    completed replicas), `try_num` is never decremented by the Ab script (so the loop ends when `sample_number`
    is exhausted), with `similarity_search` one row is written and `sample_number` drops to 0, without it every
    replica consumes one unit.  `select(tokens, h_list, l_list) -> index`.  Returns (rows [(h, l)], tokens, passes)."""
    tok = Tokenizer()
    rows, passes = [], 0
    B = tokens.shape[0]
    tokens = tokens.clone()
    order = torch.from_numpy(np.asarray(loc).copy())
    while sample_number > 0 and try_number > 0:
        tokens = model.sample(tokens.to(device), region.to(device), chain.to(device), order, rng=rng, seed=seed,
                              seq_offset=passes * B).cpu()
        passes += 1
        h, l = split_and_untokenize(tokens, tok)
        if similarity_search:
            k = select(tokens, h, l)
            rows.append((h[k], l[k]))
            sample_number = 0
        else:
            for g_h, g_l in zip(h, l):
                if sample_number == 0:
                    break
                rows.append((g_h, g_l))
                sample_number -= 1
    return rows, tokens, passes


def nano_sample_passes(model, tokens, region, loc, sample_number: int, try_number: int, validate=None, rng="philox",
                       seed=0, device="cuda", log=lambda *_: None):
    """nanosample.py:316-353: every replica is checked with `validate` (the reference numbers it with
    abnumber.Chain); a valid one is written and consumes one unit of `sample_number`, an invalid one is written only
    on the last try; `try_num` drops by one per replica looked at, valid or not; tokens are not re-masked between
    passes.  Returns (written sequences, passes)."""
    tok = Tokenizer()
    rows, passes = [], 0
    B = tokens.shape[0]
    tokens = tokens.clone()
    order = torch.from_numpy(np.asarray(loc).copy())
    while sample_number > 0 and try_number > 0:
        tokens = model.sample(tokens.to(device), region.to(device), None, order, rng=rng, seed=seed,
                              seq_offset=passes * B).cpu()
        passes += 1
        for g_h in [tok.idx2seq(s) for s in tokens]:
            if sample_number == 0:
                break
            log(g_h)
            try:
                if validate is not None:
                    validate(g_h)
                rows.append(g_h)
                sample_number -= 1
            except Exception:                                     # the reference's bare except (nanosample.py:347)
                if try_number == 1:
                    rows.append(g_h)
                log("Need to re sample again.")
            try_number -= 1
    return rows, passes


def abnumber_validator():
    """`Chain(g_h, scheme='imgt')` of nanosample.py:342, or None when abnumber is not installed."""
    try:
        from abnumber import Chain
    except ImportError:
        return None
    return lambda seq: Chain(seq, scheme="imgt")


def humanize_antibody(model, mouse_h: str, mouse_l: str, batch_size: int = 1, finetune: bool = True,
                      pad_region: int = 0, number: NumberFn = anarci_numbering, shuffle: bool = True,
                      rng: str = "philox", seed: int = 0, device="cuda", sample_method: str = "FR",
                      graft: GraftFn = abnumber_graft_chain, sample_number: int = 1, try_number: int = 1,
                      similarity_search: bool = True, chain_factory=None):
    """One antibody through sample.py:481-536 -> (rows [(h, l)], tokens [B, 291], parental tokens [B, 291]).

    The sequences are numbered once (the memoised `number`), the parental tokens come from the same slot grid."""
    cache: Dict[str, tuple] = {}

    def number_once(seq):
        if seq not in cache:
            cache[seq] = number(seq)
        return cache[seq]

    if sample_method == "inpaint":
        tokens, region, chain, _, loc, tok = batch_inpaint_input_element(mouse_h, mouse_l, batch_size, pad_region, graft)
    else:
        tokens, region, chain, _, loc, tok = batch_input_element(mouse_h, mouse_l, batch_size, pad_region, finetune, number_once)
    parent = tok.seq2idx(get_input_element(mouse_h, mouse_l, pad_region, number_once)[1]).unsqueeze(0).expand(batch_size, -1)
    if shuffle:
        np.random.shuffle(loc)                                   # numpy global RNG, like sample.py:498

    def select(tokens_out, h, l):
        if chain_factory is not None:
            return select_the_most_similarity_seq(mouse_h, mouse_l, h, l, chain_factory)[0]
        try:
            return select_the_most_similarity_seq(mouse_h, mouse_l, h, l)[0]
        except RuntimeError:                                     # no abnumber: aligned identity on the slot grid
            return select_most_similar(parent, tokens_out)

    rows, out, _ = ab_sample_passes(model, tokens, region, chain, loc, sample_number, try_number, similarity_search,
                                    select, rng, seed, device)
    return rows, out, parent


def humanize_nanobody(model, vhh: str, batch_size: int = 1, inpaint_sample: bool = False,
                      number: NumberFn = anarci_numbering, shuffle: bool = True, rng: str = "philox", seed: int = 0,
                      device="cuda", sample_number: int = 1, try_number: int = 10, validate="abnumber", log=lambda *_: None):
    """One VHH through nanosample.py:301-353 -> written sequences."""
    tokens, region, loc, tok = nano_batch_input_element(vhh, inpaint_sample, batch_size, number)
    if shuffle:
        np.random.shuffle(loc)                                   # nanosample.py:315
    if validate == "abnumber":
        validate = abnumber_validator()
    rows, _ = nano_sample_passes(model, tokens, region, loc, sample_number, try_number, validate, rng, seed, device, log)
    return rows


def load_checkpoint(path: str, version: str = "finetune", kind: str = "ab"):
    """Checkpoint formats of the reference: returns (config, state_dict).

    The resolved pair is cached next to the checkpoint as `<ckpt>.hudiffpack` (hudiff_b200/packcache.py: flat tensors
    keyed by the sha256 of the checkpoint), so later starts skip the un-pickling of the training state."""
    from . import packcache
    return packcache.load_cached(path, version, kind, resolve_checkpoint)


def resolve_checkpoint(path: str, version: str = "finetune", kind: str = "ab"):
    """Un-pickle a reference checkpoint and resolve its layout: (config, state_dict)."""
    from .config import install_easydict_shim
    install_easydict_shim()
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    if kind == "ab":
        config = ckpt["pretrain_config"] if version == "finetune" else ckpt["config"]
        state = ckpt["model"]
    elif version == "finetune_vh":                             # nanofinetune.py:531-539
        config = {"name": "nano", "model": ckpt["infilling_params"]}
        state = {k.partition("infilling_pretrain.")[2]: v for k, v in ckpt["model"].items()
                 if k.startswith("infilling_pretrain.")}
    else:
        config, state = ckpt["config"], ckpt["model"]
    state = {(k[7:] if k.startswith("module.") else k): v for k, v in state.items()}
    return config, state


# --------------------------------------------------------------------------------------------- #
# single-antibody front-ends (sample_for_anti_cdr.py, sample_for_nano_cdr.py)
# --------------------------------------------------------------------------------------------- #
def read_fasta(path: str):
    """[(description, sequence)] of a FASTA file (what Bio.SeqIO.parse yields, without Biopython)."""
    recs, desc, seq = [], None, []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                if desc is not None:
                    recs.append((desc, "".join(seq)))
                desc, seq = line[1:], []
            else:
                seq.append(line)
    if desc is not None:
        recs.append((desc, "".join(seq)))
    return recs


def h_l_from_fasta(path: str):
    """Heavy and light chain of a complex FASTA: records whose description contains 'heavy chain' /
    'light chain' (sample_for_anti_cdr.py:50-67)."""
    heavy = light = None
    for desc, seq in read_fasta(path):
        if "heavy chain" in desc:
            heavy = seq
        elif "light chain" in desc:
            light = seq
    assert heavy is not None and light is not None, "Reading the fasta has problem."
    return heavy, light


def nano_from_fasta(path: str):
    """The record whose description contains 'Nanobody' (sample_for_nano_cdr.py:33-46)."""
    nano = None
    for desc, seq in read_fasta(path):
        if "Nanobody" in desc:
            nano = seq
    assert nano is not None, "Reading the fasta has problem."
    return nano


def sample_until(model, tokens, region, chain, loc, sample_number: int, untokenize, rng="philox", seed=0,
                 device="cuda", validate=None):
    """The `while sample_number > 0` loop of the single-antibody scripts
    (sample_for_anti_cdr.py:177-214, sample_for_nano_cdr.py:165-201).

    Every pass runs the whole denoising loop on the B replicas; like the reference the tokens are NOT
    re-masked between passes (later passes re-sample every slot of the already complete sequences), every
    replica consumes one unit of `sample_number`, and only sequences not seen before are returned.
    `untokenize(tokens) -> list of hashable results`; `validate(result)` may raise to reject a sample
    (the nanobody script numbers the sample with abnumber first)."""
    seen, unique = set(), []
    passes = 0
    tokens = tokens.clone()
    while sample_number > 0:
        tokens = model.sample(tokens.to(device), region.to(device), None if chain is None else chain.to(device),
                              torch.from_numpy(np.asarray(loc).copy()), rng=rng, seed=seed + passes).cpu()
        passes += 1
        for res in untokenize(tokens):
            if sample_number == 0:
                break
            if res not in seen:
                if validate is not None:
                    validate(res)
                seen.add(res)
                unique.append(res)
            sample_number -= 1
    return unique, passes
