"""Batch humanization of paired antibodies on B200 -- drop-in for the reference entry point
`antibody_scripts/sample.py` (same flags, same result CSV / FASTA formats; sample.py:379-584).

Differences, all additive: the denoising loop runs on the GPU through `model.sample`
(`hudiff_b200`), `--numbering_json` supplies pre-computed IMGT numbering when ANARCI/abnumber are
not installed, `--rng torch_cpu_replay` reproduces the reference's CPU random stream token for
token, and without abnumber the best replica is chosen by slot-grid identity.
"""
import argparse
import os
import sys

sys.path.append(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import pandas as pd
import torch

from hudiff_b200 import model_selected, seed_all
from hudiff_b200 import sampling as S


def write_fasta(path, names, hseqs, lseqs):
    """abnumber Chain.to_fasta layout: '>name description' then the sequence (sample.py:32-54)."""
    with open(path, "w") as f:
        for name, h, l in zip(names, hseqs, lseqs):
            f.write(f">{name} VH\n{h}\n>{name} VL\n{l}\n")


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--ckpt", type=str, default="checkpoints/antibody/hudiffab.pt")
    p.add_argument("--ckpt_version", type=str, default="finetune", choices=["pretrain", "finetune"])
    p.add_argument("--data_fpath", type=str, default="humanization_pair_data_filter.csv")
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--sample_number", type=int, default=1)
    p.add_argument("--try_number", type=int, default=1)
    p.add_argument("--seed", type=int, default=2023)
    p.add_argument("--sample_order", type=str, default="shuffle")
    p.add_argument("--sample_method", type=str, default="FR", choices=["FR", "inpaint"])
    p.add_argument("--similarity_search", type=bool, default=True)
    p.add_argument("--fa_version", type=str, default="v007")
    # additions
    p.add_argument("--numbering_json", type=str, default=None)
    p.add_argument("--graft_json", type=str, default=None)
    p.add_argument("--rng", type=str, default="philox", choices=["philox", "torch_cpu_replay"])
    p.add_argument("--precision", type=str, default="fp16", choices=["fp16", "fp16x3"])
    p.add_argument("--device", type=str, default="cuda:0")
    p.add_argument("--out_dir", type=str, default=None)
    args = p.parse_args(argv)

    seed_all(args.seed)                                          # before model construction, like the reference
    config, state = S.load_checkpoint(args.ckpt, args.ckpt_version, "ab")
    finetune = args.ckpt_version == "finetune"
    model = model_selected(config)
    model.load_state_dict(state)
    model = model.to(args.device).eval().set_precision(args.precision)
    n_region = (config["model"] if isinstance(config, dict) else config.model)["n_region"]
    pad_region = 7 if n_region > 7 else 0                      # sample.py:462-465
    number = S.FixtureNumbering.from_json(args.numbering_json) if args.numbering_json else S.anarci_numbering
    graft = S.FixtureGraft.from_json(args.graft_json) if args.graft_json else S.abnumber_graft_chain

    out_dir = args.out_dir or os.path.dirname(os.path.dirname(os.path.abspath(args.ckpt)))
    os.makedirs(out_dir, exist_ok=True)
    save_fpath = os.path.join(out_dir, "sample_humanization_result.csv")
    with open(save_fpath, "w", encoding="UTF-8") as f:
        f.write("Specific,name,hseq,lseq,\n")
    df = pd.read_csv(args.data_fpath)
    mouse_df = df[df["type"] == "mouse"]
    names, hs, ls = [], [], []
    for idx, line in enumerate(mouse_df.itertuples()):
        with open(save_fpath, "a", encoding="UTF-8") as f:
            f.write(f"mouse,{line.name},{line.h_seq},{line.l_seq}\n")
        # the `while sample_number > 0 and try_num > 0` loop of sample.py:499-536 (S.ab_sample_passes)
        rows, _, _ = S.humanize_antibody(
            model, line.h_seq, line.l_seq, args.batch_size, finetune, pad_region, number,
            shuffle=args.sample_order == "shuffle", rng=args.rng, seed=args.seed + idx, device=args.device,
            sample_method=args.sample_method, graft=graft, sample_number=args.sample_number,
            try_number=args.try_number, similarity_search=args.similarity_search)
        with open(save_fpath, "a", encoding="UTF-8") as f:
            for g_h, g_l in rows:
                f.write(f"humanization,{line.name}human_sample,{g_h},{g_l}\n")
                names.append(f"{args.fa_version}human{len(names)}")
                hs.append(g_h)
                ls.append(g_l)
    write_fasta(os.path.join(out_dir, "sample_identity.fa"), names, hs, ls)
    return save_fpath


if __name__ == "__main__":
    main()
