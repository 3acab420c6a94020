"""Ad-hoc GPU bring-up script: run a few golden cases and print diffs vs the oracle goldens."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sushie_b200
from sushie_b200 import config

g = np.load(os.path.join(ROOT, "tests/golden/synthetic_small.npz"))
gs = np.load(os.path.join(ROOT, "tests/golden/example_ss.npz"))
cases = {"k1": (1, dict(L=5)), "k2": (2, dict(L=5)), "k3": (3, dict(L=6)), "k4": (4, dict(L=4)),
         "k2rho": (2, dict(L=4, no_update=True, rho=[0.5])), "k2noupd": (2, dict(L=4, no_update=True))}
teams = [int(t) for t in (sys.argv[1:] or ["1"])]
for team in teams:
    config.team_size = team
    for tag, (k, kw) in cases.items():
        Xs = [g[f"{tag}_X{i}"].astype(float) for i in range(k)]
        ys = [g[f"{tag}_y{i}"].copy() for i in range(k)]
        t0 = time.time()
        try:
            res = sushie_b200.infer_sushie(Xs, ys, min_tol=1e-4, **kw)
        except Exception:
            traceback.print_exc()
            continue
        dt = time.time() - t0
        n_ref = int(g[f"{tag}_n_iter"])
        e_ref = g[f"{tag}_elbo"]
        m = min(len(res.elbo), len(e_ref))
        with np.errstate(all="ignore"):
            de = np.nanmax(np.abs(res.elbo[1:m] - e_ref[1:m]) / np.abs(e_ref[1:m])) if m > 1 else float("nan")
        dp = np.max(np.abs(res.pip_all - g[f"{tag}_pip_all"]))
        print(f"team={team} ind {tag}: n_iter={len(res.elbo)-1} (ref {n_ref}) inc={res.elbo_increase} rel dELBO={de:.2e} max dPIP={dp:.2e} first elbo={res.elbo[1:4]} ref={e_ref[1:4]} t={dt:.3f}s", flush=True)
    for tag in ("k2", "k3"):
        t0 = time.time()
        try:
            res = sushie_b200.infer_sushie_ss(gs[f"{tag}_lds"], gs[f"{tag}_ns"], gs[f"{tag}_zs"], L=10, min_tol=1e-3)
        except Exception:
            traceback.print_exc()
            continue
        dt = time.time() - t0
        e_ref = gs[f"{tag}_tol3_elbo"]
        m = min(len(res.elbo), len(e_ref))
        de = np.nanmax(np.abs(res.elbo[1:m] - e_ref[1:m]) / np.abs(e_ref[1:m]))
        dp = np.max(np.abs(res.pip_all - gs[f"{tag}_tol3_pip_all"]))
        print(f"team={team} ss {tag}: n_iter={len(res.elbo)-1} (ref {int(gs[f'{tag}_tol3_n_iter'])}) rel dELBO={de:.2e} max dPIP={dp:.2e} first={res.elbo[1:4]} ref={e_ref[1:4]} t={dt:.3f}s", flush=True)
