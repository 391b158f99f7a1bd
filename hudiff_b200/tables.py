"""IMGT slot tables that define HuDiff's fixed-length inputs and sampling masks.

These are the constants of the reference's pre-processing; values must be equal
for the B200 path to be a drop-in.  They are generated here from a run-length
description instead of literal lists.

Reference:
  dataset/preprocess.py:195-212  HEAVY_POSITIONS_dict   (152 slots)
  dataset/preprocess.py:294-309  LIGHT_POSITIONS_dict   (139 slots)
  dataset/preprocess.py:214-223  HEAVY_CDR_INDEX        (pretrain mask: ==0 is sampled)
  dataset/preprocess.py:252-265  HEAVY_CDR_KABAT_NO_VERNIER (finetune mask)
  dataset/preprocess.py:269-277  INPAINT_HEAVY_CDR_INDEX (Nb --inpaint_sample)
  dataset/preprocess.py:311-320  LIGHT_CDR_INDEX
  dataset/preprocess.py:348-362  LIGHT_CDR_KABAT_NO_VERNIER
  dataset/oas_pair_dataset_new.py:25-40  HEAVY_REGION_INDEX / LIGHT_REGION_INDEX
`tests/test_tables.py` checks every table against the reference when it is present.
"""
from __future__ import annotations

import string

H_LEN = 152
L_LEN = 139
AB_LEN = H_LEN + L_LEN


def _expand(runs):
    out = []
    for value, count in runs:
        out.extend([value] * count)
    return out


def _imgt_slots(n_insert: int, last: int):
    """IMGT position labels in slot order: 1..111, 111A.., ..112A, 112..last."""
    letters = string.ascii_uppercase[:n_insert]
    labels = [str(i) for i in range(1, 112)]
    labels += [f"111{c}" for c in letters]
    labels += [f"112{c}" for c in reversed(letters)]
    labels += [str(i) for i in range(112, last + 1)]
    return {lab: i for i, lab in enumerate(labels)}


HEAVY_POSITIONS_dict = _imgt_slots(12, 128)
LIGHT_POSITIONS_dict = _imgt_slots(6, 127)

HEAVY_CDR_INDEX = _expand([(0, 26), (1, 12), (0, 17), (2, 10), (0, 39), (3, 37), (0, 11)])
LIGHT_CDR_INDEX = _expand([(0, 26), (1, 12), (0, 17), (2, 10), (0, 39), (3, 25), (0, 10)])

HEAVY_CDR_KABAT_NO_VERNIER = _expand(
    [(0, 26), (1, 14), (0, 14), (2, 20), (0, 30), (3, 37), (0, 9), (4, 2)])
LIGHT_CDR_KABAT_NO_VERNIER = _expand(
    [(0, 23), (1, 17), (0, 11), (5, 4), (2, 14), (0, 35), (3, 25), (0, 9), (4, 1)])

INPAINT_HEAVY_CDR_INDEX = _expand(
    [(0, 26), (1, 12), (0, 3), (4, 1), (0, 6), (4, 2), (0, 1), (4, 1), (0, 2), (2, 12),
     (0, 38), (3, 37), (0, 11)])

HEAVY_REGION_INDEX = _expand([(0, 26), (1, 12), (2, 17), (3, 10), (4, 39), (5, 37), (6, 11)])
LIGHT_REGION_INDEX = _expand([(0, 26), (1, 12), (2, 17), (3, 10), (4, 39), (5, 25), (6, 10)])

assert len(HEAVY_POSITIONS_dict) == len(HEAVY_CDR_INDEX) == len(HEAVY_REGION_INDEX) == H_LEN
assert len(LIGHT_POSITIONS_dict) == len(LIGHT_CDR_INDEX) == len(LIGHT_REGION_INDEX) == L_LEN
assert len(HEAVY_CDR_KABAT_NO_VERNIER) == len(INPAINT_HEAVY_CDR_INDEX) == H_LEN
assert len(LIGHT_CDR_KABAT_NO_VERNIER) == L_LEN
