#!/usr/bin/env python
"""Config 5 of BASELINE.json, scaled to what one call can hold: a genome-wide style batch of genes with
ragged widths, each fitted once on all samples and once per cross-validation fold (1 + 5 fits per gene),
sharded over the GPUs of the box by the host work queue (``sushie_b200.batch``).

  python scripts/bench_cv.py [--genes G] [--devices 0 1 ...] [--min-tol 1e-4]

Genes are config-3 genes (``bench.make_gene``: AR(1) haplotypes, hard calls, h2 = 0.2) with
``p_g ~ U{500..4000}`` drawn from ``default_rng(20263000 + g)``; folds come from the CLI's own
``_prepare_cv`` (threefry row shuffle, 5 chunks, rank-normalised phenotypes), so a fold fit trains on
400 rows per ancestry.  Two numbers per run, one JSON line:

  * ``resident``: all 6·G problems in one device-resident batch, CUDA-event time of the IBSS launches;
  * ``api``: wall time of ``infer_sushie_batch`` (host prologue, upload, IBSS, download, credible sets)
    plus the out-of-fold prediction / r-squared step of ``cli._cv_scores`` -- the work `--cv` does.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bench  # noqa: E402  (gene generator of the headline workload)

SEED_P = 20263000
CV_NUM = 5


def gene_width(g: int, lo: int, hi: int) -> int:
    return int(np.random.default_rng(SEED_P + g).integers(lo, hi + 1))


def _gen(args):
    g, lo, hi = args
    return bench.make_gene(g, bench.N_IND, gene_width(g, lo, hi))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=200)
    ap.add_argument("--p-min", type=int, default=500)
    ap.add_argument("--p-max", type=int, default=4000)
    ap.add_argument("--devices", type=int, nargs="*", default=[0])
    ap.add_argument("--max-iter", type=int, default=500)
    ap.add_argument("--min-tol", type=float, default=1e-4)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--no-api", action="store_true")
    ap.add_argument("--no-resident", action="store_true")
    args = ap.parse_args()

    import multiprocessing as mp

    t0 = time.time()
    with mp.get_context("fork").Pool(os.cpu_count() or 1) as pool:
        genes = pool.map(_gen, [(g, args.p_min, args.p_max) for g in range(args.genes)], chunksize=2)
    t_gen = time.time() - t0

    from sushie_b200 import _capi, cli, config, infer  # noqa: E402

    # ---- the fits of every gene: full data + CV_NUM training folds (cli.py:323-419 of the reference)
    t0 = time.time()
    specs, folds_of = [], []
    for Xs, ys in genes:
        specs.append(dict(Xs=Xs, ys=ys))
        folds = cli._prepare_cv(Xs, ys, CV_NUM, 12345)
        folds_of.append(folds)
        for f in folds:
            specs.append(dict(Xs=f.train_geno, ys=f.train_pheno))
    t_folds = time.time() - t0
    n_fits = len(specs)
    widths = [s["Xs"][0].shape[1] for s in specs]
    cells = float(sum(sum(x.shape[0] for x in s["Xs"]) * s["Xs"][0].shape[1] for s in specs))

    common = dict(L=bench.L_EFF, max_iter=args.max_iter, min_tol=args.min_tol)

    # ---- resident arm (first device): one batch, CUDA-event time of the IBSS launches
    dev = args.devices[0]
    eng = _capi.get_engine(dev)
    preps = [infer._prepare_kw(**dict(common, **s)) for s in specs]
    opts = eng.make_opts(args.max_iter, args.min_tol, _capi.SB_I8, config.team_size, config.single_phase)
    resident = None
    if args.no_resident:
        preps = []
    for b in ([eng.create_batch([p.problem for p in preps], opts)] if preps else []):
        for _ in range(args.warmup):
            b.run()
        ms, n_ser, n_iter, alg = 0.0, 0, 0, 0.0
        for _ in range(args.steps):
            st = b.run()
            ms += st["device_ms"]
            n_ser += st["n_ser"]
            n_iter += st["n_iter"]
            alg += st["algorithmic_bytes"]
        peak, peak_src = bench.hbm_peak()
        sec = ms / 1e3
        resident = {
            "fits_per_sec": n_fits * args.steps / sec, "genes_per_sec": args.genes * args.steps / sec,
            "ms_per_step": ms / args.steps, "mean_iters_per_fit": n_iter / (n_fits * args.steps),
            "ser_per_sec": n_ser / sec, "launches_per_step": st["n_run_launches"],
        }
        # SURVEY.md 8(d) algorithmic bytes: one traversal of each fit's genotype tensor per single-effect update at
        # fp64 (the stats count the stored bytes, 1 B per genotype)
        resident["fp64_equivalent_GBps"] = 8.0 * alg / sec / 1e9
        resident["frac_of_hbm_peak_fp64_equivalent"] = resident["fp64_equivalent_GBps"] / peak
        resident["stored_GBps"] = alg / sec / 1e9
        resident["peak"] = peak
        resident["peak_source"] = peak_src
        b.close()
    del preps

    # ---- public API arm: wall clock of what `finemap --cv` does per gene, all genes in one call
    api = None
    if not args.no_api:
        for s in range(2):
            t0 = time.perf_counter()
            fits = infer.infer_sushie_batch([dict(common, **sp) for sp in specs], devices=args.devices)
            t_fit = time.perf_counter() - t0
            t1 = time.perf_counter()
            r2 = []
            for gi in range(args.genes):
                base = gi * (1 + CV_NUM)
                r2.append(cli._cv_scores(folds_of[gi], fits[base + 1: base + 1 + CV_NUM]))
            t_cv = time.perf_counter() - t1
        api = {
            "fits_per_sec": n_fits / (t_fit + t_cv), "genes_per_sec": args.genes / (t_fit + t_cv),
            "wall_s_fits": t_fit, "wall_s_cv_scores": t_cv, "devices": args.devices,
            "chunk_loci": config.chunk_loci or "auto (memory budget %.0f GiB per chunk at fp64 size; with several GPU "
                                               "workers at least one chunk per worker)" % (config.chunk_bytes / 2**30),
            "mean_cv_adj_r2_ancestry0": float(np.mean([r[0][0] for r in r2])),
            "n_credible_sets": int(sum(len(set(f.cs["CSIndex"])) if len(f.cs) else 0 for f in fits[:: 1 + CV_NUM])),
        }

    line = {
        "bench": "config 5 (scaled): genes x (1 full + 5 CV-fold fits), ragged widths, host work queue",
        "genes": args.genes, "fits": n_fits, "cv_num": CV_NUM, "k": bench.K_ANC, "n_full": bench.N_IND,
        "p_range": [args.p_min, args.p_max], "p_mean": float(np.mean(widths)), "L": bench.L_EFF,
        "max_iter": args.max_iter, "min_tol": args.min_tol, "genotype_storage": "i8",
        "genotype_cells_per_pass": cells, "datagen_s": t_gen, "fold_split_s": t_folds,
        "resident": resident, "api": api, "steps": args.steps, "warmup": args.warmup,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
