#!/usr/bin/env python
"""Benchmark of the IBSS hot path on BASELINE.json's headline workload.

Workload (config 3): a batch of synthetic 3-ancestry individual-level loci,
n = 500 per ancestry, p = 2000 SNPs, L = 10, fitted to convergence with the Python-API
defaults (max_iter = 500, min_tol = 1e-4).  One "step" = one pass of the whole batch
through the IBSS engine.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--genes B]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference        # the numpy restatement of the reference on host cores

Prints ONE JSON line (rank 0).  `value` = loci/s with the genotypes already resident in HBM
(device time by CUDA events on the engine's stream, max over ranks); `e2e` = loci/s through
the C ABI with host buffers (int8 genotypes in pinned memory in, full posterior tables out),
copies inside the timed region.  The genotypes of this workload are hard calls (0/1/2), so
the default device storage is one byte per entry (`--storage i8`: centring / scaling folded
into the vectors, all arithmetic fp64); `roofline` refers to that kernel with the stored
bytes as algorithmic bytes.  `roofline_dense_fp64` (N = 1) repeats the resident run with the
standardised fp64 matrix (`--storage f64`), the configuration where DRAM traffic equals the
algorithmic bytes and the HBM-roofline claim is made.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_ANC, N_IND, P_SNP, L_EFF = 3, 500, 2000, 10
RHO_ANC = (0.90, 0.80, 0.85)
SEED0 = 20260000
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


# ----------------------------------------------------------------------------------------
# synthetic loci


def make_gene(g: int, n: int = N_IND, p: int = P_SNP):
    """Gene g: AR(1)-latent haplotypes -> 0/1/2 genotypes (int8), y = X_std b + e at h2 = 0.2."""
    from scipy.signal import lfilter
    from scipy.stats import norm

    rng = np.random.default_rng(SEED0 + g)
    Xs, ys = [], []
    maf = rng.uniform(0.05, 0.5, size=p)
    thr = norm.ppf(1.0 - maf)
    n_causal = g % 4
    causal = rng.choice(p, size=n_causal, replace=False) if n_causal else np.zeros(0, dtype=int)
    cov = 0.8 * np.ones((K_ANC, K_ANC)) + 0.2 * np.eye(K_ANC)
    b = rng.multivariate_normal(np.zeros(K_ANC), cov, size=n_causal) if n_causal else np.zeros((0, K_ANC))
    for a in range(K_ANC):
        rho = RHO_ANC[a]
        c = math.sqrt(1.0 - rho * rho)
        geno = np.zeros((n, p), dtype=np.int8)
        for _ in range(2):
            e = rng.standard_normal((n, p))
            e[:, 0] /= c
            z = lfilter([c], [1.0, -rho], e, axis=1)
            geno += (z > thr).astype(np.int8)
        sd = geno.std(axis=0)
        for j in np.nonzero(sd == 0)[0]:
            geno[0, j] = 1 if geno[0, j] == 0 else geno[0, j] - 1
        Xs.append(geno)
        if n_causal:
            gf = geno[:, causal].astype(np.float64)
            gstd = (gf - gf.mean(0)) / gf.std(0)
            gv = gstd @ b[:, a]
            s2e = np.var(gv) * (1.0 / 0.2 - 1.0)
            ys.append(gv + rng.standard_normal(n) * math.sqrt(s2e))
        else:
            ys.append(rng.standard_normal(n))
    return Xs, ys


def _gen_into(args):
    g, n, p = args
    return make_gene(g, n, p)


def generate_genes(first: int, count: int, workers: int, n=N_IND, p=P_SNP):
    import multiprocessing as mp

    jobs = [(first + i, n, p) for i in range(count)]
    if workers <= 1 or count < 4:
        return [make_gene(*j) for j in jobs]
    with mp.get_context("fork").Pool(workers) as pool:
        return pool.map(_gen_into, jobs, chunksize=max(1, count // (workers * 4)))


# ----------------------------------------------------------------------------------------
# CPU legs (the only places allowed to execute oracle/)


def _oracle_fit(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)
    except Exception:
        pass
    from oracle import ibss_fast

    g, max_iter, min_tol, n, p, keep = args
    Xs, ys = make_gene(g, n, p)
    t0 = time.perf_counter()
    fit = ibss_fast.fit_individual([x.astype(np.float64) for x in Xs], ys, L=L_EFF, max_iter=max_iter, min_tol=min_tol)
    out = {"seconds": time.perf_counter() - t0, "n_iter": fit.n_iter}
    if keep:  # what the in-bench parity check compares (effects in IBSS order, before the re-ordering)
        out.update(alpha=fit.alpha, effect_covar=fit.effect_covar, elbo_last=float(fit.elbo[-1]))
    return out


def cpu_pass(first_gene: int, n_loci: int, cores: int, max_iter: int, min_tol: float, n=N_IND, p=P_SNP, keep=False):
    """Fit n_loci genes with the numpy port of the reference (oracle/ibss_fast.py: same quantities as the literal
    restatement oracle/ibss.py, two traversals of X per single-effect update), one single-threaded process per core.
    Returns (wall seconds, total iterations, per-locus records)."""
    import multiprocessing as mp

    jobs = [(first_gene + i, max_iter, min_tol, n, p, keep) for i in range(n_loci)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(min(cores, n_loci)) as pool:
        out = pool.map(_oracle_fit, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    return wall, int(sum(o["n_iter"] for o in out)), out


def credible_members(alpha, threshold=0.95, max_size=250):
    """Per effect: the SNPs of the `threshold` credible set (descending alpha, cumulative cut, infer.py:903-915), or
    None when the set is larger than max_size (a null effect: its membership is decided by rounding)."""
    out = []
    for a in alpha:
        order = np.argsort(-a, kind="stable")
        size = int(np.count_nonzero(np.cumsum(a[order]) < threshold)) + 1
        out.append(frozenset(order[:size].tolist()) if size <= max_size else None)
    return out


def parity_record(gpu_fits, cpu_records):
    """GPU fits vs the CPU port on the SAME loci (the cpu_baseline sample), effect by effect in IBSS order:
    north_star tolerances |dPIP| <= 1e-4, identical credible-set membership, prior covariance within 1e-4 relative."""
    n_same_iter, d_pip, d_cov, d_elbo, sets_total, sets_same = 0, 0.0, 0.0, 0.0, 0, 0
    for fit, rec in zip(gpu_fits, cpu_records):
        n_same_iter += int(fit.n_iter == rec["n_iter"])
        pip_g = -np.expm1(np.log1p(-fit.alpha).sum(axis=0))
        pip_c = -np.expm1(np.log1p(-rec["alpha"]).sum(axis=0))
        d_pip = max(d_pip, float(np.max(np.abs(pip_g - pip_c))))
        for cg, cc in zip(fit.effect_covar, rec["effect_covar"]):
            d_cov = max(d_cov, float(np.linalg.norm(cg - cc) / np.linalg.norm(cc)))
        d_elbo = max(d_elbo, abs(float(fit.elbo[-1]) - rec["elbo_last"]) / abs(rec["elbo_last"]))
        for sg, sc in zip(credible_members(fit.alpha), credible_members(rec["alpha"])):
            if sg is None and sc is None:
                continue
            sets_total += 1
            sets_same += int(sg == sc)
    n = len(cpu_records)
    ok = n_same_iter == n and d_pip <= 1e-4 and d_cov <= 1e-4 and sets_same == sets_total
    return {
        "loci_compared": n, "same_n_iter": n_same_iter, "max_abs_dpip": d_pip, "max_rel_dcov": d_cov,
        "max_rel_delbo": d_elbo, "credible_sets_compared": sets_total, "credible_sets_identical": sets_same,
        "within_tolerance": bool(ok),
        "tolerance": "same n_iter; |dPIP| <= 1e-4; identical 95% credible-set membership (sets <= 250 SNPs); "
                     "|dC|/|C| <= 1e-4 per effect",
        "against": "oracle/ibss_fast.py on the cpu_baseline loci (numpy port; held to the literal restatement and, "
                   "through tests/refshim, to the reference's own source by the test-suite)",
    }


# ----------------------------------------------------------------------------------------
# config 5 (BASELINE.json): genes x (1 full + 5 cross-validation fold) fits through the host work queue

SEED_P = 20263000
CV_NUM = 5


def gene_width(g: int, lo: int = 500, hi: int = 4000) -> int:
    return int(np.random.default_rng(SEED_P + g).integers(lo, hi + 1))


def _gen_to_dir(args):
    g, path = args
    Xs, ys = make_gene(g, N_IND, gene_width(g))
    np.save(os.path.join(path, f"g{g}_X.npy"), np.stack(Xs))
    np.save(os.path.join(path, f"g{g}_y.npy"), np.stack(ys))
    return g


def config5_dir(world: int) -> str:
    tag = os.environ.get("MASTER_PORT", str(os.getpid())) if world > 1 else str(os.getpid())
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    return os.path.join(base, f"sushie_b200_c5_{tag}")


def config5_generate(n_genes: int, rank: int, world: int, workers: int) -> str:
    """This rank's share of the genes (g = rank mod world) into a directory every rank of the box can read: any
    rank may pull any gene from the queue.  Host work, before CUDA exists in this process (fork pool)."""
    import multiprocessing as mp

    path = config5_dir(world)
    os.makedirs(path, exist_ok=True)
    mine = [(g, path) for g in range(n_genes) if g % world == rank]
    with mp.get_context("fork").Pool(max(1, workers)) as pool:
        pool.map(_gen_to_dir, mine, chunksize=4)
    return path


def config5_run(args, path, rank, local_rank, world, dist):
    """Strong scaling: a FIXED total of genes, one process per GPU, ONE global cost-sorted chunk list pulled through
    an atomic counter in the torch.distributed store; every gene = the public-API work of `finemap --cv` (fold split,
    6 fits, credible sets, tables, out-of-fold r2).  NCCL only gathers the per-gene result records."""
    import shutil

    import torch

    import sushie_b200
    from sushie_b200 import batch, cli

    n_genes = args.config5_genes
    widths = [gene_width(g) for g in range(n_genes)]
    kw = dict(L=L_EFF, max_iter=args.max_iter, min_tol=args.min_tol, devices=[local_rank])

    def process(ids):
        specs, folds_of = [], []
        for g in ids:
            X = np.load(os.path.join(path, f"g{g}_X.npy"), mmap_mode="r")
            y = np.load(os.path.join(path, f"g{g}_y.npy"))
            Xs, ys = [X[a] for a in range(K_ANC)], [y[a] for a in range(K_ANC)]
            folds = cli._prepare_cv(list(Xs), list(ys), CV_NUM, 12345)
            folds_of.append(folds)
            specs.append(dict(Xs=Xs, ys=ys))
            specs.extend(dict(Xs=f.train_geno, ys=f.train_pheno, lazy_tables=True) for f in folds)
        fits = sushie_b200.infer_sushie_batch(specs, **kw)
        out = []
        for i, g in enumerate(ids):
            base = i * (1 + CV_NUM)
            full = fits[base]
            r2 = cli._cv_scores(folds_of[i], fits[base + 1: base + 1 + CV_NUM])
            out.append((len(full.elbo) - 1, int(len(set(full.cs.CSIndex.values.tolist()))), float(np.sum(full.pip_all)),
                        [float(r[0]) for r in r2]))
        return out

    from sushie_b200 import config as sb_config

    inner_workers, sb_config.workers_per_device = sb_config.workers_per_device, 1  # the queue's two workers overlap already
    store = dist.distributed_c10d._get_default_store() if world > 1 else None
    if world > 1:
        dist.barrier()  # every rank's genes are on disk
    # warm-up: a few genes per rank through a private queue with the same worker count (imports, kernels, and the
    # workers' engines -- stream, events, page-locked staging -- which a long-lived process keeps between batches)
    warm = [g for g in range(n_genes) if g % world == rank][: 2 * args.c5_workers]
    batch.run_distributed([float(widths[g]) for g in warm], lambda ids: process([warm[i] for i in ids]),
                          workers_per_rank=args.c5_workers, min_items=2, max_items=2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    records, trace = batch.run_distributed(
        [float(w) for w in widths], process, store=store, rank=rank, world=world, workers_per_rank=args.c5_workers,
        key=f"sushie_b200/config5/{os.environ.get('MASTER_PORT', '0')}", min_items=args.c5_min_items,
        max_items=args.c5_max_items)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    sb_config.workers_per_device = inner_workers
    if world > 1:
        t = torch.tensor([wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # (also the closing barrier of the timed region)
        wall = float(t[0])
        gathered = [None] * world
        dist.all_gather_object(gathered, (records, trace))  # result records only; inputs never move
    else:
        gathered = [(records, trace)]
    out = None
    if rank == 0:
        recs = dict(pair for part, _ in gathered for pair in part)
        assert sorted(recs) == list(range(n_genes)), "every gene exactly once"
        per_rank = [sum(n for _, _, n in tr) for _, tr in gathered]
        out = {
            "workload": f"config 5: {n_genes} genes x (1 full + {CV_NUM} cross-validation fold fits), p ~ U{{500..4000}}, "
                        f"n = {N_IND} (folds {N_IND - N_IND // CV_NUM}) per ancestry, k = {K_ANC}, L = {L_EFF}, to convergence",
            "genes": n_genes, "fits": n_genes * (1 + CV_NUM), "scaling": "strong (fixed total)", "n_gpus": world,
            "wall_s": wall, "genes_per_sec": n_genes / wall, "fits_per_sec": n_genes * (1 + CV_NUM) / wall,
            "path": "per gene: load from the shared host directory, threefry fold split (folds = row views of the gene's "
                    "matrix, uploaded once), infer_sushie_batch (host prologue, upload, IBSS, device-side set selection, "
                    "download in effect order, purity, PIPs; the full fit's tables are built, the folds' are not asked "
                    f"for), out-of-fold r2 -- one process per GPU, {args.c5_workers} worker threads each taking turns in "
                    "the IBSS kernel, chunks pulled from one global cost-sorted list (sushie_b200.batch.StoreQueue)",
            "workers_per_rank": args.c5_workers, "chunk_genes": [args.c5_min_items, args.c5_max_items],
            "collective": "none on the data path; NCCL all_gather of the per-gene result records at the end",
            "chunks": sum(len(tr) for _, tr in gathered), "genes_per_rank": per_rank,
            "mean_iters_full_fit": float(np.mean([r[0] for r in recs.values()])),
            "credible_sets": int(sum(r[1] for r in recs.values())),
            "mean_cv_adj_r2": [float(np.mean([r[3][a] for r in recs.values()])) for a in range(K_ANC)],
            "data": "synthetic genes generated before the timed region into a host directory shared by the ranks",
        }
    if world > 1:
        dist.barrier()
    if rank == 0:
        shutil.rmtree(path, ignore_errors=True)
    return out


# ----------------------------------------------------------------------------------------
# helpers


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="sb_clocks_", suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=self.fh, stderr=subprocess.DEVNULL,
            )
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(which: str):
    """DRAM read+write bytes per single-effect update of an IBSS kernel, from the newest committed `ncu --set full`
    summary of the same kernel (profiles/rNN_ncu_full_<ind_i8|ind_f64|ss_f32>.json)."""
    import glob

    name = which if which.startswith("ss_") else f"ind_{which}"
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_ncu_full_{name}.json")), reverse=True):
        try:
            with open(path) as fh:
                return float(json.load(fh)["dram_bytes_per_ser"]), os.path.relpath(path, ROOT)
        except Exception:
            continue
    return None, None


def prep_y(ys):
    out, rv = [], []
    for y in ys:
        y = y - y.mean()
        y = y / y.std()
        out.append(y)
        rv.append(np.var(y, ddof=1))
    return out, rv


# ----------------------------------------------------------------------------------------


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = max(1, args.ref_loci or cores)
    times, iters = [], 0
    for _ in range(args.warmup):  # warm-up passes are bounded: 2 loci x 5 iterations (imports, page-in)
        cpu_pass(0, min(2, sample), cores, 5, 0.0)
    for _ in range(args.steps):
        wall, it, _ = cpu_pass(0, sample, cores, args.max_iter, args.min_tol)
        times.append(wall)
        iters += it
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    line = {
        "impl": "reference",
        "metric": "loci_per_sec",
        "value": value,
        "unit": "loci/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {
            "value": value, "unit": "loci/s", "cores": min(cores, sample), "kind": "port",
            "sample": f"each step: the first {sample} loci of the workload fitted to convergence, one single-threaded "
                      f"process per core (numpy port oracle/ibss_fast.py; the reference's JAX path is not "
                      f"installable: no jax/equinox wheels)",
        },
        "e2e": {"value": value, "unit": "loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "iters_per_sec": iters / (sum(times)),
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    """The workload, identical for both arms (`--impl reference` fits a bounded sample of it per step, described in
    its `cpu_baseline.sample`)."""
    return {
        "workload": "config 3: synthetic 3-ancestry individual-level loci (AR(1) haplotypes, h2=0.2, 0-3 causal)",
        "k": K_ANC, "n_per_ancestry": N_IND, "p": P_SNP, "L": L_EFF,
        "genes_per_gpu": args.genes, "max_iter": args.max_iter, "min_tol": args.min_tol,
        "stop_rule": ("to convergence (Python API defaults)" if (args.max_iter, args.min_tol) == (500, 1e-4) else
                      f"fixed {args.max_iter} iterations (min_tol = 0 disables the stop, infer.py:537)" if args.min_tol == 0
                      else f"to convergence (max_iter = {args.max_iter}, min_tol = {args.min_tol})"),
        "l2": "inputs_larger_than_L2 (>= %d MB of genotypes per locus resident, batch of %d loci >> 126 MB)" % (
            K_ANC * N_IND * P_SNP // 1000000, args.genes),
        "parallelism": f"loci sharded over {args.gpus} GPU(s), no data-path collective",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genes", type=int, default=1000, help="loci per GPU per step")
    ap.add_argument("--max-iter", type=int, default=500)
    ap.add_argument("--min-tol", type=float, default=1e-4)
    ap.add_argument("--storage", default="b2", choices=["b2", "i8", "f64", "f32"],
                    help="device storage of the genotypes: b2 = hard calls 0/1/2 as bit planes (products by table "
                         "look-up), i8 = raw hard calls (1 B/entry), both with centring/scaling folded into the "
                         "vectors; f64 / f32 = standardised dense matrix; arithmetic is fp64 in every case")
    ap.add_argument("--team", type=int, default=0)
    ap.add_argument("--single-phase", action="store_true")
    ap.add_argument("--ref-loci", type=int, default=0)
    ap.add_argument("--cpu-loci", type=int, default=0, help="loci in the cpu_baseline sample (0 = one per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-dense-arm", action="store_true", help="skip the extra fp64-storage roofline run (N=1 only)")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--dense-steps", type=int, default=3)
    ap.add_argument("--no-ss-arm", action="store_true", help="skip the summary-level (config 4) roofline run (N=1 only)")
    ap.add_argument("--ss-p", type=int, default=10000)
    ap.add_argument("--ss-iters", type=int, default=20)
    ap.add_argument("--config5-genes", type=int, default=2368,
                    help="genes of the config-5 arm (fixed total for every N: strong scaling; 0 = skip)")
    ap.add_argument("--c5-workers", type=int, default=3, help="config-5 arm: worker threads per rank pulling chunks")
    ap.add_argument("--c5-min-items", type=int, default=24, help="config-5 arm: smallest chunk (genes)")
    ap.add_argument("--c5-max-items", type=int, default=64, help="config-5 arm: largest chunk (genes)")
    ap.add_argument("--no-api-arm", action="store_true", help="skip the public-API end-to-end run")
    ap.add_argument("--api-genes", type=int, default=0, help="loci of the public-API end-to-end run (0 = --genes)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    B = args.genes

    # ---- host work before CUDA exists in this process: data generation, CPU baseline
    gen_workers = max(1, cores // max(1, world))
    t0 = time.time()
    genes = generate_genes(rank * B, B, gen_workers)
    c5_path = config5_generate(args.config5_genes, rank, world, gen_workers) if args.config5_genes > 0 else None
    t_gen = time.time() - t0

    cpu_baseline, cpu_records = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_loci or min(cores, B)
        wall, it, cpu_records = cpu_pass(0, sample, cores, args.max_iter, args.min_tol, keep=True)
        cpu_baseline = {
            "value": sample / wall, "unit": "loci/s", "cores": min(cores, sample), "kind": "port",
            "sample": f"first {sample} loci of the workload to convergence, numpy port (oracle/ibss_fast.py), one "
                      f"single-threaded process per core, {wall:.1f} s wall",
            "iters_per_sec": it / wall,
        }

    import torch
    import torch.distributed as dist

    # one process per GPU: keep the BLAS / OpenMP pools of the ranks inside their share of the host cores (8 ranks x
    # 32 spinning BLAS threads on a 32-core box halved the per-rank throughput of the config-5 arm)
    blas_limit = None
    if world > 1:
        try:
            from threadpoolctl import threadpool_limits

            blas_limit = threadpool_limits(limits=max(1, cores // world))
        except Exception:
            pass
        torch.set_num_threads(max(1, cores // world))
        # ... and the engine's download-copy helper threads (read at their first use)
        os.environ.setdefault("SUSHIE_B200_COPY_THREADS", str(max(1, min(4, cores // world // 2))))

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from sushie_b200 import _capi
    from sushie_b200._common import build_priors

    eng = _capi.Engine(local_rank)
    effect_covar, adj = build_priors(K_ANC, L_EFF, None, None, False, summary=False)

    # pinned host buffers: int8 genotypes in, posterior tables out
    x_pin = torch.empty((B, K_ANC, N_IND, P_SNP), dtype=torch.int8).pin_memory()
    x_np = x_pin.numpy()
    for i, (Xs, _) in enumerate(genes):
        for a in range(K_ANC):
            x_np[i, a] = Xs[a]
    ys_all = [g[1] for g in genes]
    del genes

    def build_problems():
        probs = []
        for i in range(B):
            ys, rv = prep_y(ys_all[i])
            probs.append(_capi.IndProblem(
                [x_np[i, a] for a in range(K_ANC)], ys, L_EFF, None, rv, effect_covar, adj.times, adj.plus,
                em_update=True, standardize=_capi.SB_CENTER | _capi.SB_SCALE))
        return probs

    storage = {"b2": _capi.SB_B2, "i8": _capi.SB_I8, "f64": _capi.SB_F64, "f32": _capi.SB_F32}[args.storage]
    elem_bytes = {"b2": 0.5, "i8": 1, "f64": 8, "f32": 4}[args.storage]
    opts = eng.make_opts(args.max_iter, args.min_tol, storage, args.team, args.single_phase)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident arm: upload + standardise once, then W warm-up and K timed passes
    probs = build_problems()
    batch = eng.create_batch(probs, opts)
    for _ in range(args.warmup):
        st = batch.run()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms, n_ser, n_iter, alg_bytes, launches = 0.0, 0, 0, 0.0, 0
    for _ in range(args.steps):
        st = batch.run()
        dev_ms += st["device_ms"]
        n_ser += st["n_ser"]
        n_iter += st["n_iter"]
        alg_bytes += st["algorithmic_bytes"]
        launches += st["n_run_launches"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    clocks = sampler.stop()
    fits = batch.fetch()
    pip_sum = float(sum(np.sum(-np.expm1(np.log1p(-f.alpha).sum(axis=0))) for f in fits[:8]))
    # the GPU fits of the loci the CPU port has just fitted in this very run: compared, not just printed
    parity = parity_record(fits[: len(cpu_records)], cpu_records) if cpu_records else None
    n_nonfinite = int(sum(f.status != 0 for f in fits))
    geom = {k: st[k] for k in ("team_size", "n_teams", "rows_per_band", "n_stages", "smem_bytes")}
    batch.close()
    del fits

    # ---- dense fp64 arm (N=1): the same batch with the standardised fp64 matrix resident, where DRAM
    # traffic == algorithmic bytes; this is the run the ">= 60 % of the HBM roofline" claim is made on
    dense = None
    if args.storage != "f64" and world == 1 and not args.no_dense_arm:
        d_opts = eng.make_opts(args.max_iter, args.min_tol, _capi.SB_F64, args.team, args.single_phase)
        d_batch = eng.create_batch(build_problems(), d_opts)
        d_batch.run()
        d_sampler = ClockSampler(local_rank)
        d_sampler.start()
        d_ms, d_alg, d_ser, d_steps = 0.0, 0.0, 0, max(3, args.dense_steps)
        for _ in range(d_steps):
            d_st = d_batch.run()
            d_ms += d_st["device_ms"]
            d_alg += d_st["algorithmic_bytes"]
            d_ser += d_st["n_ser"]
        d_clocks = d_sampler.stop()
        d_batch.close()
        d_ach = d_alg / (d_ms / 1e3) / 1e9
        d_traffic, d_src = ncu_traffic("f64")
        dense = {
            "kernel": "sb::ibss_kernel<3,double,false,1>", "bound": "hbm", "achieved": d_ach, "unit": "GB/s",
            "loci_per_sec": B * d_steps / (d_ms / 1e3), "ms_per_step": d_ms / d_steps,
            "algorithmic_bytes_per_ser": K_ANC * N_IND * P_SNP * 8,
            "traffic": d_traffic * d_ser / d_steps if d_traffic else None, "traffic_source": d_src,
            "steps": d_steps, "warmup": 1, "clocks": d_clocks,
        }

    # ---- summary-level arm (N=1): BASELINE.json config 4, one k=3, p=10 000 locus, LD stored fp32 (1.2 GB read per
    # single-effect update), L=10, EM on, fixed iterations, co-operative team over the whole device
    ss_arm = None
    if world == 1 and not args.no_ss_arm:
        import scripts.bench_ss as bench_ss

        lds, ns, zs, _ = bench_ss.make_locus(args.ss_p, np.float32)
        ec, adj_ss = build_priors(K_ANC, L_EFF, None, None, False, summary=True)
        prob = _capi.SsProblem(lds, ns.reshape(-1), zs, L_EFF, None, [1.0] * K_ANC, ec, adj_ss.times, adj_ss.plus, True)
        s_opts = eng.make_opts(args.ss_iters, 0.0, _capi.SB_F32, 0, True)
        s_batch = eng.create_batch([prob], s_opts)
        del lds
        for _ in range(3):
            s_batch.run()
        s_sampler = ClockSampler(local_rank)
        s_sampler.start()
        s_ms, s_ser, s_steps = 0.0, 0, 5
        for _ in range(s_steps):
            s_st = s_batch.run()
            s_ms += s_st["device_ms"]
            s_ser += s_st["n_ser"]
        s_clocks = s_sampler.stop()
        s_batch.close()
        s_bytes = float(K_ANC) * args.ss_p * args.ss_p * 4
        s_traffic, s_src = ncu_traffic("ss_f32")
        ss_arm = {
            "workload": f"config 4: summary-level, k={K_ANC}, p={args.ss_p}, L={L_EFF}, fp32 LD storage, EM on, "
                        f"{args.ss_iters} fixed iterations",
            "kernel": "sb::ibss_kernel<3,float,true,0>", "team_size": s_st["team_size"], "bound": "hbm",
            "achieved": s_ser * s_bytes / (s_ms / 1e3) / 1e9, "unit": "GB/s", "us_per_ser": 1e3 * s_ms / s_ser,
            "ms_per_iteration": s_ms / s_ser * L_EFF, "algorithmic_bytes_per_ser": s_bytes,
            "traffic": s_traffic * s_ser / s_steps if s_traffic else None, "traffic_source": s_src,
            "steps": s_steps, "warmup": 3, "clocks": s_clocks,
            "l2": "1.2 GB of LD per single-effect update >> 126 MB L2",
        }

    # ---- end-to-end arm: host buffers through the C ABI, copies in the timed region
    e2e = None
    if not args.no_e2e:
        import ctypes as C

        sizes = [L_EFF * P_SNP, L_EFF * P_SNP * K_ANC, L_EFF * P_SNP * K_ANC * K_ANC, L_EFF * K_ANC * K_ANC,
                 L_EFF, L_EFF * P_SNP, L_EFF * K_ANC * K_ANC, K_ANC, args.max_iter + 1]
        names = ["alpha", "post_mean", "post_mean_sq", "weighted_sum_covar", "kl", "log_bf", "effect_covar",
                 "resid_var", "elbo"]
        per = sum(sizes)
        out_pin = torch.empty((B, per), dtype=torch.float64).pin_memory()
        out_np = out_pin.numpy()
        res = (_capi.SbResult * B)()
        pd_t = C.POINTER(C.c_double)
        for i in range(B):
            off = 0
            for nm, sz in zip(names, sizes):
                setattr(res[i], nm, out_np[i, off:off + sz].ctypes.data_as(pd_t))
                off += sz
        e2e_ms = []
        for s in range(1 + args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            probs = build_problems()  # host prologue (phenotype centring/scaling, priors)
            arr = (_capi.SbIndProblem * B)()
            for i, pr in enumerate(probs):
                pr.fill(arr[i])
            stats = _capi.SbRunStats()
            eng._check(eng.lib.sb_ibss_ind(eng._ctx, arr, B, C.byref(opts), res, C.byref(stats)), "sb_ibss_ind")
            barrier()
            if s >= 1:
                e2e_ms.append(1e3 * (time.perf_counter() - t0))
        e2e_step = float(np.mean(e2e_ms))
        h2d = B * (K_ANC * N_IND * P_SNP + 8 * (K_ANC * N_IND + 2 * K_ANC + L_EFF * K_ANC * K_ANC + 2 * K_ANC * K_ANC))
        d2h = B * 8 * (per - (args.max_iter + 1)) + 8 * int(stats.n_iter) + 8 * B
        e2e = {"ms": e2e_step, "h2d": h2d, "d2h": d2h, "launches": int(stats.n_launches)}

    # ---- public-API arm: numpy arrays in -> list of SushieResult out (host prologue, upload, IBSS, download, effect
    # re-ordering, credible sets + purity, output tables) -- what a user of `infer_sushie` gets, per locus
    api = None
    if not args.no_api_arm:
        import sushie_b200

        n_api = min(B, args.api_genes or B)
        api_probs = [dict(Xs=[x_np[i, a] for a in range(K_ANC)], ys=ys_all[i]) for i in range(n_api)]
        kw = dict(L=L_EFF, max_iter=args.max_iter, min_tol=args.min_tol, devices=[local_rank])
        sushie_b200.infer_sushie_batch(api_probs[: min(n_api, 64)], **kw)  # warm-up (imports, pandas, shuffle cache)
        barrier()
        t0 = time.perf_counter()
        results = sushie_b200.infer_sushie_batch(api_probs, **kw)
        # (both tables of every result exist when the call returns; reading them is inside the timed region anyway)
        n_cs = int(sum(len(set(r.cs.CSIndex.values.tolist())) for r in results))
        n_alpha_cols = int(sum(r.alphas.shape[1] for r in results))
        barrier()
        api = {"ms": 1e3 * (time.perf_counter() - t0), "loci": n_api, "n_cs": n_cs, "alphas_columns": n_alpha_cols}
        del results

    # ---- config 5 arm: the multi-GPU design itself (work queue, ragged genes, CV folds, public API)
    config5 = config5_run(args, c5_path, rank, local_rank, world, dist) if c5_path else None

    # ---- reduce over ranks (max time), gather a small result record with NCCL
    if world > 1:
        t = torch.tensor([dev_ms, wall_ms, e2e["ms"] if e2e else 0.0, api["ms"] if api else 0.0], device="cuda",
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(t[0]), float(t[1])
        if e2e:
            e2e["ms"] = float(t[2])
        if api:
            api["ms"] = float(t[3])
        cnt = torch.tensor([n_ser, n_iter, alg_bytes, pip_sum], device="cuda", dtype=torch.float64)
        gathered = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(gathered, cnt)
        n_ser = int(sum(float(g[0]) for g in gathered))
        n_iter = int(sum(float(g[1]) for g in gathered))
        alg_bytes = float(sum(float(g[2]) for g in gathered))

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        total_loci = B * world
        value = total_loci / (ms_per_step / 1e3)
        peak, peak_src = hbm_peak()
        # per-GPU achieved bandwidth of the single persistent launch
        achieved = (alg_bytes / world) / (dev_ms / 1e3) / 1e9
        traffic_per_ser, traffic_src = ncu_traffic(args.storage)
        # algorithmic bytes = one traversal of the genotype tensor per single-effect update at the
        # precision under test (fp64 arithmetic: k*n*p*8 B; fp32 storage run: *4).  Hard-call storage keeps the same
        # fp64 computation but reads 1 B per genotype, so its fraction is > 1 by compression -- stated here with the
        # stored-bytes figure and the ncu DRAM traffic next to it; the dense arm below is the <= 1 roofline run.
        alg_elem = 4 if args.storage == "f32" else 8
        ser_per_gpu_s = (n_ser / world) / (dev_ms / 1e3)
        achieved_alg = ser_per_gpu_s * K_ANC * N_IND * P_SNP * alg_elem / 1e9
        roofline = {
            "bound": "hbm", "achieved": achieved_alg, "peak": peak, "unit": "GB/s", "frac": achieved_alg / peak,
            "traffic": (traffic_per_ser * (n_ser / world) / args.steps) if traffic_per_ser else None,
            "traffic_source": (f"{traffic_src}: ncu dram read+write bytes per single-effect update "
                               f"({traffic_per_ser:.0f} B) x updates per step (one persistent launch "
                               f"plus its short tail phases)") if traffic_per_ser else None,
            "peak_source": peak_src,
            "algorithmic_bytes_per_ser": K_ANC * N_IND * P_SNP * alg_elem,
            "stored_bytes_per_ser": K_ANC * N_IND * P_SNP * elem_bytes,
            "achieved_stored_bytes": achieved, "frac_stored_bytes": achieved / peak,
            # DRAM view: ncu read+write bytes per update x updates / time -- what the memory system actually moves
            "achieved_dram": (traffic_per_ser * ser_per_gpu_s / 1e9) if traffic_per_ser else None,
            "frac_dram": (traffic_per_ser * ser_per_gpu_s / 1e9 / peak) if traffic_per_ser else None,
            "limiter": {"b2": "shared-memory table look-ups / issue", "i8": "fp64 issue"}.get(args.storage, "hbm"),
            "note": ("achieved = #single-effect updates x k*n*p*%d B (one traversal of the genotype tensor at the "
                     "precision under test) / CUDA-event time, per GPU" % alg_elem)
                    + ("; genotypes are stored as hard calls (bit planes or bytes: compressed), so `frac` counts fp64-equivalent "
                       "bytes and is NOT a fraction of HBM bandwidth: the memory system moves `achieved_dram` "
                       "(`frac_dram` of peak) and the kernel is bound by fp64 / issue slots (`limiter`); the "
                       "HBM-roofline claims are roofline_dense_fp64 and roofline_ss_f32, where DRAM traffic equals "
                       "the algorithmic bytes" if args.storage in ("i8", "b2") else ""),
        }
        line = {
            "metric": "loci_per_sec",
            "value": value,
            "unit": "loci/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args),
            "iters_per_sec": n_iter / (dev_ms / 1e3),
            "ser_per_sec": n_ser / (dev_ms / 1e3),
            "mean_iters_per_locus": n_iter / (total_loci * args.steps),
            "wall_ms_per_step": wall_ms / args.steps,
            "gpu_launches": launches,
            "kernel": dict(geom, name="sb::ibss_kernel<3,%s>" % {"b2": "uint8_t,false,0", "i8": "int8_t,false,1", "f64": "double,false,1",
                                                         "f32": "float,false,0"}[args.storage],
                           launches_per_step=launches / args.steps, genotype_storage=args.storage,
                           stored_mb_per_locus=K_ANC * N_IND * P_SNP * elem_bytes / 1e6),
            "roofline": roofline,
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                       "samples": clocks["samples"]},
            "datagen_s": t_gen,
            "check_pip_sum_first8": pip_sum,
        }
        if e2e:
            line["e2e"] = {
                "value": total_loci / (e2e["ms"] / 1e3), "unit": "loci/s",
                "h2d_bytes_per_step": e2e["h2d"] * world, "d2h_bytes_per_step": e2e["d2h"] * world,
                "ms_per_step": e2e["ms"], "gpu_launches_per_step": e2e["launches"],
                "path": "sb_ibss_ind (C ABI): pinned int8 genotypes -> device column statistics -> IBSS -> posterior tables to pinned host",
            }
        if api:
            line["e2e_api"] = {
                "value": api["loci"] * world / (api["ms"] / 1e3), "unit": "loci/s", "ms_per_step": api["ms"],
                "loci_per_gpu": api["loci"], "steps": 1, "warmup": 1, "credible_sets_rank0": api["n_cs"],
                "path": "sushie_b200.infer_sushie_batch: numpy genotypes / phenotypes in -> SushieResult (priors, "
                        "posteriors, PIPs, credible-set and alpha tables) out, one call for the whole batch",
            }
        if dense:
            dense["peak"] = peak
            dense["frac"] = dense["achieved"] / peak
            line["roofline_dense_fp64"] = dense
        if ss_arm:
            ss_arm["peak"] = peak
            ss_arm["frac"] = ss_arm["achieved"] / peak
            line["roofline_ss_f32"] = ss_arm
        if config5:
            line["config5"] = config5
        if parity:
            line["parity"] = parity
        line["nonfinite_loci_rank0"] = n_nonfinite
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
