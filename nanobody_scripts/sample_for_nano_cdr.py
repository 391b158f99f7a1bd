"""Humanize ONE nanobody given as a FASTA on B200 -- drop-in for the reference entry point
`nanobody_scripts/sample_for_nano_cdr.py` (same flags, same result CSV / FASTA; :49-219).

Same loop semantics as the reference (`while sample_number > 0`, duplicates dropped but counted, no
re-masking between passes); each pass is one `model.sample` call.  A sample is validated by numbering it
with abnumber when that package is installed (:193), otherwise the check is skipped.  Additive flags:
`--numbering_json`, `--rng`, `--precision`, `--device`, `--out_dir`; `--seed` is honoured.
"""
import argparse
import os
import sys

sys.path.append(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from hudiff_b200 import model_selected, seed_all
from hudiff_b200 import sampling as S


def _abnumber_check(seq):
    try:
        from abnumber import Chain
    except ImportError:                                  # not installed offline: nothing to validate with
        return
    Chain(seq, scheme="imgt")


def main(argv=None):
    p = argparse.ArgumentParser(description="This program is designed to humanize non-human nanobodies.")
    p.add_argument("--ckpt", type=str, default=None)
    p.add_argument("--nano_complex_fasta", type=str, default=None)
    p.add_argument("--batch_size", type=int, default=10)
    p.add_argument("--sample_number", type=int, default=100)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--sample_order", type=str, default="shuffle")
    p.add_argument("--sample_method", type=str, default="gen", choices=["gen", "rl_gen"])
    p.add_argument("--length_limit", type=str, default="not_equal")
    p.add_argument("--model", type=str, default="finetune_vh", choices=["pretrain", "finetune_vh"])
    p.add_argument("--fa_version", type=str, default="v_nano")
    p.add_argument("--inpaint_sample", type=eval, default=True)
    p.add_argument("--structure", type=eval, default=False)
    # additions
    p.add_argument("--numbering_json", type=str, default=None)
    p.add_argument("--rng", type=str, default="philox", choices=["philox", "torch_cpu_replay"])
    p.add_argument("--precision", type=str, default="fp16", choices=["fp16", "fp16x3"])
    p.add_argument("--device", type=str, default="cuda:0")
    p.add_argument("--out_dir", type=str, default=None)
    args = p.parse_args(argv)

    seed_all(args.seed)
    pdb_name = os.path.basename(args.nano_complex_fasta).split(".")[0]
    log_dir = args.out_dir or os.path.join(os.path.dirname(args.nano_complex_fasta), f"{pdb_name}_{args.model}_vhh")
    os.makedirs(log_dir, exist_ok=True)

    config, state = S.load_checkpoint(args.ckpt, args.model, "nb")
    model = model_selected(config)
    model.load_state_dict(state)
    model = model.to(args.device).eval().set_precision(args.precision)
    number = S.FixtureNumbering.from_json(args.numbering_json) if args.numbering_json else S.anarci_numbering

    nano_chain = S.nano_from_fasta(args.nano_complex_fasta)
    save_fpath = os.path.join(log_dir, "sample_humanization_result.csv")
    with open(save_fpath, "w", encoding="UTF-8") as f:
        f.write("Specific,name,hseq,\n")
        f.write(f"Nano,{pdb_name},{nano_chain}\n")

    tokens, region, loc, tok = S.nano_batch_input_element(nano_chain, args.inpaint_sample, args.batch_size, number)
    if args.sample_order == "shuffle":
        np.random.shuffle(loc)
    unique, passes = S.sample_until(model, tokens, region, None, loc, args.sample_number,
                                    lambda t: [tok.idx2seq(s) for s in t], rng=args.rng, seed=args.seed,
                                    device=args.device, validate=_abnumber_check)
    with open(save_fpath, "a", encoding="UTF-8") as f:
        for g_h in unique:
            f.write(f"humanization,{pdb_name},{g_h}\n")
    with open(os.path.join(log_dir, "sample_identity.fa"), "w") as f:      # for BioPhi OASis, like nanosample
        for i, g_h in enumerate(unique):
            f.write(f">{args.fa_version}human{i}\n{g_h}\n")
    if args.structure:                                                     # one FASTA per sample (structure prediction)
        for i, g_h in enumerate(unique):
            with open(os.path.join(log_dir, f"{pdb_name}_human_{i}.fasta"), "w") as f:
                f.write(f">{pdb_name}_human_{i}\n{g_h}\n")
    print(f"{len(unique)} unique humanized sequences of {args.sample_number} samples in {passes} passes -> {save_fpath}")
    return save_fpath


if __name__ == "__main__":
    main()
