"""Anchors against artefacts the reference authors produced with their own code and shipped in
``/root/reference/data`` (``data/make_example.py``): the per-SNP GWAS z-scores (``*.gwas``, scipy
``linregress`` on standardised genotypes) and the LD matrices (``*.ld``, rounded to 4 digits).  Our own VCF /
plink decoders + standardisation must reproduce them from the raw genotype and phenotype files.  Runs only
where the reference tree is present (this container), never on the GPU box."""
import os

import numpy as np
import pandas as pd
import pytest

from sushie_b200 import io

DATA = "/root/reference/data"
pytestmark = pytest.mark.skipif(not os.path.isdir(DATA), reason="reference data tree not present")


@pytest.mark.parametrize("anc", ["EUR", "AFR", "EAS"])
def test_decoders_reproduce_the_shipped_gwas_and_ld(anc):
    bim, fam, bed = io.read_vcf(f"{DATA}/vcf/{anc}.vcf")
    bim2, fam2, bed2 = io.read_triplet(f"{DATA}/plink/{anc}")
    np.testing.assert_array_equal(bed, bed2)
    pheno = pd.read_csv(f"{DATA}/{anc}.pheno", sep="\t", header=None, dtype={0: object})
    assert (pheno[0].values == fam.iid.values).all()
    gwas = pd.read_csv(f"{DATA}/{anc}.gwas", sep="\t")
    ld = io.read_ld(f"{DATA}/{anc}.ld")
    cols = [int(np.nonzero(bim.snp.values == s)[0][0]) for s in gwas.snp]
    # the shipped files are in the first ancestry's allele coding: flip the columns whose counted allele differs
    g = bed[:, cols].copy()
    flip = bim.a1.values[cols] != gwas.a1.values
    assert (bim.a0.values[cols][flip] == gwas.a1.values[flip]).all()
    g[:, flip] = 2 - g[:, flip]
    gs = (g - g.mean(axis=0)) / g.std(axis=0)
    y = pheno[1].values.astype(float)
    ys = (y - y.mean()) / y.std()
    n = len(ys)
    beta = gs.T @ ys / n
    resid = ys[:, None] - gs * beta
    se = np.sqrt((resid ** 2).sum(axis=0) / (n - 2) / n)
    np.testing.assert_allclose(beta / se, gwas.z.values, rtol=1e-6, atol=1e-7)  # shipped values carry fp32-level noise
    np.testing.assert_allclose(beta, gwas.beta.values, rtol=1e-6, atol=1e-8)
    assert list(ld.columns) == list(gwas.snp)
    np.testing.assert_allclose(np.round(gs.T @ gs / n, 4), ld.values, atol=1.01e-4)
