"""Batch humanization of nanobodies (VHH) on B200 -- drop-in for `nanobody_scripts/nanosample.py`
(flags and CSV layout of nanosample.py:195-368).  The loop runs on the GPU via `model.sample`."""
import argparse
import os
import sys

sys.path.append(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import pandas as pd

from hudiff_b200 import model_selected, seed_all
from hudiff_b200 import sampling as S


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--ckpt", type=str, default="nanofinetune.pt")
    p.add_argument("--data_fpath", type=str, default="abnativ_select_vhh.csv")
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--sample_number", type=int, default=1)
    p.add_argument("--try_number", type=int, default=10)
    p.add_argument("--seed", type=int, default=2023)
    p.add_argument("--sample_order", type=str, default="shuffle")
    p.add_argument("--model", type=str, default="finetune_vh", choices=["pretrain", "finetune_vh"])
    p.add_argument("--fa_version", type=str, default="v_nano")
    p.add_argument("--inpaint_sample", type=eval, default=False)
    # additions
    p.add_argument("--numbering_json", type=str, default=None)
    p.add_argument("--rng", type=str, default="philox", choices=["philox", "torch_cpu_replay"])
    p.add_argument("--precision", type=str, default="fp16", choices=["fp16", "fp16x3"])
    p.add_argument("--device", type=str, default="cuda:0")
    p.add_argument("--out_dir", type=str, default=None)
    args = p.parse_args(argv)

    seed_all(args.seed)
    config, state = S.load_checkpoint(args.ckpt, args.model, "nb")
    model = model_selected(config)
    model.load_state_dict(state)
    model = model.to(args.device).eval().set_precision(args.precision)
    number = S.FixtureNumbering.from_json(args.numbering_json) if args.numbering_json else S.anarci_numbering

    out_dir = args.out_dir or os.path.dirname(os.path.dirname(os.path.abspath(args.ckpt)))
    os.makedirs(out_dir, exist_ok=True)
    save_fpath = os.path.join(out_dir, "sample_humanization_result.csv")
    with open(save_fpath, "w", encoding="UTF-8") as f:
        f.write("Specific,name,hseq,\n")
    df = pd.read_csv(args.data_fpath)
    fasta = []
    for idx, line in enumerate(df.itertuples()):
        with open(save_fpath, "a", encoding="UTF-8") as f:
            f.write(f"nano,{idx},{line.vhhseq}\n")
        # retry / validation loop of nanosample.py:316-353 (S.nano_sample_passes): abnumber.Chain when installed
        seqs = S.humanize_nanobody(model, line.vhhseq, args.batch_size, args.inpaint_sample, number,
                                   shuffle=args.sample_order == "shuffle", rng=args.rng, seed=args.seed + idx,
                                   device=args.device, sample_number=args.sample_number, try_number=args.try_number)
        with open(save_fpath, "a", encoding="UTF-8") as f:
            for g_h in seqs:
                f.write(f"humanization,{idx}human_sample,{g_h}\n")
                fasta.append((f"{args.fa_version}human{len(fasta)}", g_h))
    with open(os.path.join(out_dir, "sample_identity.fa"), "w") as f:
        for name, seq in fasta:
            f.write(f">{name}\n{seq}\n")
    return save_fpath


if __name__ == "__main__":
    main()
