"""Humanize ONE antibody given as a complex FASTA (or two sequences) on B200 -- drop-in for the reference
entry point `antibody_scripts/sample_for_anti_cdr.py` (same flags, same result CSV; :72-224).

The `while sample_number > 0` loop keeps the reference's behaviour: B replicas per pass, no re-masking
between passes, duplicates are dropped but still count.  Each pass is one `model.sample` call (the whole
T-step loop on the GPU).  Additive flags: `--numbering_json` (IMGT numbering fixture when ANARCI / abnumber
are not installed), `--rng`, `--precision`, `--device`, `--out_dir`; `--seed` is honoured (the reference
parses it but leaves `seed_all` commented out, :114).
"""
import argparse
import os
import sys

sys.path.append(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from hudiff_b200 import model_selected, seed_all
from hudiff_b200 import sampling as S


def main(argv=None):
    p = argparse.ArgumentParser(description="This program is designed to humanize non-human antibodies.")
    p.add_argument("--ckpt", type=str, default="checkpoints/antibody/hudiffab.pt")
    p.add_argument("--anti_complex_fasta", type=str, default=None)
    p.add_argument("--heavy_seq", type=str)
    p.add_argument("--light_seq", type=str)
    p.add_argument("--log_dirpath", type=str, default="antibody_sample_log/")
    p.add_argument("--batch_size", type=int, default=10)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--sample_number", type=int, default=10)
    p.add_argument("--sample_order", type=str, default="shuffle")
    p.add_argument("--sample_type", type=str, default="pair")
    p.add_argument("--finetune", type=str, default=True)
    # additions
    p.add_argument("--numbering_json", type=str, default=None)
    p.add_argument("--rng", type=str, default="philox", choices=["philox", "torch_cpu_replay"])
    p.add_argument("--precision", type=str, default="fp16", choices=["fp16", "fp16x3"])
    p.add_argument("--device", type=str, default="cuda:0")
    p.add_argument("--out_dir", type=str, default=None)
    args = p.parse_args(argv)

    seed_all(args.seed)
    if args.anti_complex_fasta is not None:
        mouse_heavy, mouse_light = S.h_l_from_fasta(args.anti_complex_fasta)
        pdb_name = os.path.basename(args.anti_complex_fasta).split(".")[0]
    else:
        mouse_heavy, mouse_light, pdb_name = args.heavy_seq, args.light_seq, "Unkown"
    tag = f"{pdb_name}_{args.sample_order}_{args.sample_type}"
    log_dir = args.out_dir or os.path.join(args.log_dirpath or os.path.dirname(args.anti_complex_fasta), tag)
    os.makedirs(log_dir, exist_ok=True)

    config, state = S.load_checkpoint(args.ckpt, "finetune", "ab")       # ckpt['pretrain_config'] / ckpt['model'] (:143-147)
    model = model_selected(config)
    model.load_state_dict(state)
    model = model.to(args.device).eval().set_precision(args.precision)
    number = S.FixtureNumbering.from_json(args.numbering_json) if args.numbering_json else S.anarci_numbering
    finetune = bool(args.finetune)                                       # any non-empty string is truthy, as in the reference

    save_fpath = os.path.join(log_dir, "sample_humanization_result.csv")
    with open(save_fpath, "w", encoding="UTF-8") as f:
        f.write("Specific,name,hseq,lseq,\n")
        f.write(f"mouse,{pdb_name},{mouse_heavy},{mouse_light}\n")

    tokens, region, chain, _, loc, tok = S.batch_input_element(mouse_heavy, mouse_light, args.batch_size, 0, finetune, number)
    if args.sample_order == "shuffle":
        np.random.shuffle(loc)
    unique, passes = S.sample_until(
        model, tokens, region, chain, loc, args.sample_number,
        lambda t: list(zip(*S.split_and_untokenize(t, tok))), rng=args.rng, seed=args.seed, device=args.device)
    with open(save_fpath, "a", encoding="UTF-8") as f:
        for g_h, g_l in unique:
            f.write(f"humanization,{pdb_name}human_sample,{g_h},{g_l}\n")
    print(f"{len(unique)} unique humanized sequences of {args.sample_number} samples in {passes} passes -> {save_fpath}")
    return save_fpath


if __name__ == "__main__":
    main()
