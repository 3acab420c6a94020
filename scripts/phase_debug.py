"""Compare phased vs single-phase scheduling locus by locus (bring-up helper)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sushie_b200 import _capi
from sushie_b200._common import build_priors

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
genes = bench.generate_genes(0, B, 8)
eng = _capi.Engine(0)
ec, adj = build_priors(3, 10, None, None, False, summary=False)
def probs():
    out = []
    for Xs, ys in genes:
        y2, rv = bench.prep_y(ys)
        out.append(_capi.IndProblem(Xs, y2, 10, None, rv, ec, adj.times, adj.plus, True, standardize=3))
    return out
res = {}
for name, sp, team in (("single", True, 0), ("phased", False, 0)):
    opts = eng.make_opts(500, 1e-4, _capi.SB_F64, team, sp)
    b = eng.create_batch(probs(), opts)
    st = b.run()
    res[name] = b.fetch()
    print(name, st)
    b.close()
for i in range(B):
    a, c = res["single"][i], res["phased"][i]
    m = min(len(a.elbo), len(c.elbo))
    d = np.abs(a.elbo[1:m] - c.elbo[1:m])
    first_bad = int(np.argmax(d > 1e-6 * np.abs(a.elbo[1:m]))) if np.any(d > 1e-6 * np.abs(a.elbo[1:m])) else -1
    print(i, "n_iter", a.n_iter, c.n_iter, "inc", a.elbo_increase, c.elbo_increase, "max|dELBO|", d.max() if m > 1 else 0, "first_bad_iter", first_bad,
          "last elbo", a.elbo[-1], c.elbo[-1])
