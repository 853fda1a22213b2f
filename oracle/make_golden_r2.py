"""TEST INFRASTRUCTURE ONLY -- round-2 additions to tests/golden/: full reverse-diffusion chains of the LIVE, UNMODIFIED
reference (`DDPM.sample_chain`, lightning.py:405-463) at the BENCHMARKED shapes (BASELINE configs 2 and 3, non-ragged,
T=500; config 2 also at the real ZINC checkpoint depth L=8) on 8-molecule slices, and pocket-conditioned chains (MOAD
prefix, `val_dataset` set as generate_with_pocket.py:250 does, all three cut-off graph types, T=50). Same mechanics as
oracle/make_golden.py (seeded weights verified by sha256, injected noise, oracle replay asserted against the reference in
the same run).   Run:  python -m oracle.make_golden_r2   (takes ~15 minutes of CPU)
"""
import dataclasses
import sys

import torch

from difflinker_b200 import synthetic
from oracle.make_golden import golden_chain
from oracle.ref_loader import load_reference


def main():
    torch.set_num_threads(8)
    ns = load_reference()
    S = synthetic.SPECS
    only = set(sys.argv[1:])
    jobs = []
    for gt in ("FC-10A-4A", "FC-4A", "4A"):
        pk = synthetic.WorkloadSpec(f"small_pocket_{gt}", B=2, N=70, n_min=70, l_min=5, l_max=5, F=9, L=2, T=50,
                                    seed=13, pocket=50, graph_type=gt)
        jobs.append((f"chain_small_pocket_{gt}", pk, 2, 4, 3, True))
    jobs += [("chain_cfg2_zinc_T500", S["cfg2_zinc"], 8, 0, 1, False),
             ("chain_cfg2_zinc_L8_T500", S["cfg2_zinc_L8"], 8, 0, 1, False),
             ("chain_cfg3_geom_T500", S["cfg3_geom"], 8, 0, 1, False)]
    for name, spec, nb, seed, keep, moad in jobs:
        if only and name not in only:
            continue
        print(name, flush=True)
        golden_chain(ns, name, spec, nb, seed=seed, keep_frames=keep, moad_val_dataset=moad)


if __name__ == "__main__":
    main()
