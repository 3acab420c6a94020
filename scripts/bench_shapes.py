#!/usr/bin/env python
"""Shape sweep of the individual-level IBSS kernels: microseconds per single-effect update and achieved
bandwidth for hard-call (int8) and dense (fp64) genotype storage at several locus widths.  Fixed
iterations (min_tol = 0), one locus per SM, random 0/1/2 genotypes.  One JSON line per (p, storage)."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=int, nargs="+", default=[1000, 2000, 3000, 4000, 6000, 8000])
    ap.add_argument("--n", type=int, default=500)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--L", type=int, default=10)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--loci", type=int, default=148)
    ap.add_argument("--storages", nargs="+", default=["i8", "f64"], choices=["b2", "i8", "f64"])
    ap.add_argument("--summary", action="store_true", help="summary-level sweep (banded LD, fp64 and fp32 storage)")
    args = ap.parse_args()
    from sushie_b200 import _capi
    from sushie_b200._common import build_priors
    eng = _capi.Engine(0)
    ec, adj = build_priors(args.k, args.L, None, None, False, summary=False)
    peak = 6550.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    if args.summary:
        return sweep_summary(args, eng, peak)
    for p in args.p:
        rng = np.random.default_rng(p)
        maf = rng.uniform(0.05, 0.5, size=p)
        Xs = [rng.binomial(2, maf, size=(args.n, p)).astype(np.int8) for _ in range(args.k)]
        for X in Xs:
            X[0, X.std(axis=0) == 0] = 1
        ys = [rng.standard_normal(args.n) for _ in range(args.k)]
        ys = [(y - y.mean()) / y.std() for y in ys]
        rv = [np.var(y, ddof=1) for y in ys]
        for name, storage, es in (("b2", _capi.SB_B2, 1), ("i8", _capi.SB_I8, 1), ("f64", _capi.SB_F64, 8)):
            if name not in args.storages or (name in ("i8", "b2") and p > _capi.I8_MAX_SNPS):
                continue
            probs = [_capi.IndProblem(Xs, ys, args.L, None, rv, ec, adj.times, adj.plus, True,
                                      standardize=_capi.SB_CENTER | _capi.SB_SCALE) for _ in range(args.loci)]
            try:
                b = eng.create_batch(probs, eng.make_opts(args.iters, 0.0, storage, 1, True))
            except Exception as err:
                print(json.dumps({"p": p, "storage": name, "error": str(err)[:120]}), flush=True)
                continue
            b.run()
            st = b.run()
            b.close()
            us = st["device_ms"] * 1e3 / (st["n_ser"] / min(args.loci, 148))
            print(json.dumps({
                "p": p, "n": args.n, "k": args.k, "storage": name, "us_per_update_per_sm": us,
                "loci_per_s_at_%d_iters" % args.iters: args.loci / (st["device_ms"] / 1e3),
                "stored_GBps": st["algorithmic_bytes"] / (st["device_ms"] / 1e3) / 1e9,
                "fp64_equiv_frac_of_hbm": st["n_ser"] * args.k * args.n * p * 8 / (st["device_ms"] / 1e3) / 1e9 / peak,
                "rows_per_band": st["rows_per_band"], "n_stages": st["n_stages"], "smem": st["smem_bytes"],
            }), flush=True)

def sweep_summary(args, eng, peak):
    from sushie_b200 import _capi
    from sushie_b200._common import build_priors
    ec, adj = build_priors(args.k, args.L, None, None, False, summary=True)
    for p in args.p:
        rng = np.random.default_rng(p)
        idx = np.arange(p)
        d = np.abs(idx[:, None] - idx[None, :])
        lds = [np.where(d < 20, rho ** d, 0.0) for rho in (0.9, 0.8, 0.7)[: args.k]]
        zs = [rng.standard_normal(p) for _ in range(args.k)]
        ns = np.array([5000.0, 4000.0, 3000.0][: args.k])
        n_loci = max(1, min(args.loci, int(40e9 // (args.k * p * p * 8))))
        for name, storage, es in (("f64", _capi.SB_F64, 8), ("f32", _capi.SB_F32, 4)):
            probs = [_capi.SsProblem(lds, ns, zs, args.L, None, [1.0] * args.k, ec, adj.times, adj.plus, True)
                     for _ in range(n_loci)]
            try:
                b = eng.create_batch(probs, eng.make_opts(args.iters, 0.0, storage, 1, True))
            except Exception as err:
                print(json.dumps({"p": p, "storage": name, "error": str(err)[:120]}), flush=True)
                continue
            b.run()
            st = b.run()
            b.close()
            us = st["device_ms"] * 1e3 / (st["n_ser"] / min(n_loci, 148))
            print(json.dumps({
                "mode": "summary", "p": p, "k": args.k, "storage": name, "loci": n_loci, "us_per_update_per_sm": us,
                "GBps": st["algorithmic_bytes"] / (st["device_ms"] / 1e3) / 1e9,
                "frac_of_hbm": st["algorithmic_bytes"] / (st["device_ms"] / 1e3) / 1e9 / peak,
                "rows_per_band": st["rows_per_band"], "n_stages": st["n_stages"], "smem": st["smem_bytes"],
            }), flush=True)


if __name__ == "__main__":
    main()
