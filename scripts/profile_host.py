"""Where does the HOST time of the config-5 path go?  One chunk of genes x (1 full + 5 fold fits) through
`infer_sushie_batch` in the main thread under cProfile (no worker threads, so the profile sees everything), with the
per-phase wall times of `config.trace_host`.  Development aid; run on a GPU box:

    python scripts/profile_host.py --genes 32 > gpurun_out/host_profile.txt 2>&1
"""
from __future__ import annotations

import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=32)
    ap.add_argument("--top", type=int, default=45)
    args = ap.parse_args()

    import sushie_b200
    from sushie_b200 import cli, config

    genes = []
    for g in range(args.genes):
        Xs, ys = bench.make_gene(g, bench.N_IND, bench.gene_width(g))
        genes.append((Xs, ys))

    def process(ids):
        specs, folds_of = [], []
        for g in ids:
            Xs, ys = genes[g]
            folds = cli._prepare_cv(list(Xs), list(ys), bench.CV_NUM, 12345)
            folds_of.append(folds)
            specs.append(dict(Xs=Xs, ys=ys))
            specs.extend(dict(Xs=f.train_geno, ys=f.train_pheno, lazy_tables=True) for f in folds)
        fits = sushie_b200.infer_sushie_batch(specs, L=bench.L_EFF, max_iter=500, min_tol=1e-4, devices=[0])
        out = []
        for i, g in enumerate(ids):
            base = i * (1 + bench.CV_NUM)
            full = fits[base]
            r2 = cli._cv_scores(folds_of[i], fits[base + 1: base + 1 + bench.CV_NUM])
            out.append((len(full.elbo) - 1, int(len(set(full.cs.CSIndex.values.tolist()))), float(np.sum(full.pip_all)),
                        [float(r[0]) for r in r2]))
        return out

    config.workers_per_device = 1
    process(list(range(min(4, args.genes))))  # warm-up
    config.trace_host = True
    ids = list(range(args.genes))
    t0 = time.perf_counter()
    prof = cProfile.Profile()
    prof.enable()
    process(ids)
    prof.disable()
    wall = time.perf_counter() - t0
    n_fits = len(ids) * (1 + bench.CV_NUM)
    print(f"chunk of {len(ids)} genes / {n_fits} fits: {wall * 1e3:.0f} ms wall = {wall * 1e3 / n_fits:.2f} ms per fit")
    s = io.StringIO()
    pstats.Stats(prof, stream=s).sort_stats("cumulative").print_stats(args.top)
    print(s.getvalue())
    s = io.StringIO()
    pstats.Stats(prof, stream=s).sort_stats("tottime").print_stats(args.top)
    print(s.getvalue())


if __name__ == "__main__":
    main()
