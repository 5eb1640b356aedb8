"""Synthetic haplotype data sets of the shapes named in BASELINE.json (sample count x SNP count x array density).

Sites take their minor-allele frequencies from UKBB array SNPs (data/ukbb_maf.npz, extracted from the
reference's FILES/UKBB.frq) and sit on a uniform physical grid with a 1 cM/Mb map.  Haplotypes are
Li-Stephens mosaics of a small founder panel, so that pairs share cM-scale identical tracts and the
GERMLINE-style seeding has candidates to find (independent Bernoulli haplotypes would give none).
"""
import gzip
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")


def ukbb_maf(chrom, n_sites):
    """MAFs of the first n_sites UKBB array SNPs of chromosome `chrom` (1, 20 or 22), tiled if more are asked."""
    z = np.load(os.path.join(_DATA, "ukbb_maf.npz"))
    m = z[f"chr{chrom}"]
    reps = -(-n_sites // len(m))
    return np.clip(np.tile(m, reps)[:n_sites], 1e-3, 0.5).astype(np.float64)


def make_sites(n_sites, span_bp):
    step = span_bp // n_sites
    bp = 10_000 + np.arange(n_sites, dtype=np.int64) * step
    cm = bp * 1e-6
    return bp, cm


def make_haplotypes(n_haps, maf, cm, seed, n_founders=None, switch_scale_cm=2.0, flip=1e-3, block=4096):
    """[n_haps][n_sites] uint8 alleles (1 = the allele whose frequency is `maf`)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = len(maf)
    F = n_founders or max(64, n_haps // 50)
    founders = (rng.random((F, L)) < maf[None, :]).astype(np.uint8)
    dcm = np.diff(cm, prepend=cm[0])
    p_switch = 1.0 - np.exp(-dcm / switch_scale_cm)
    p_switch[0] = 1.0
    out = np.empty((n_haps, L), np.uint8)
    for lo in range(0, n_haps, block):
        hi = min(n_haps, lo + block)
        n = hi - lo
        sw = rng.random((n, L)) < p_switch[None, :]
        seg = np.cumsum(sw, axis=1) - 1
        nseg = int(seg.max()) + 1
        choice = rng.integers(0, F, size=(n, nseg))
        fidx = np.take_along_axis(choice, seg, axis=1)
        h = founders[fidx, np.arange(L)[None, :]]
        h ^= (rng.random((n, L)) < flip).astype(np.uint8)
        out[lo:hi] = h
    return out


def write_dataset(root, haps, bp, cm, chrom=1):
    """Write <root>.hap.gz / .map / .samples in the formats FastSMC reads (Data.cpp:98-141,212-249,397-515)."""
    n_haps, L = haps.shape
    assert n_haps % 2 == 0
    os.makedirs(os.path.dirname(os.path.abspath(root)), exist_ok=True)
    with open(root + ".samples", "w") as f:
        f.write("ID_1 ID_2 missing\n0 0 0\n")
        for i in range(n_haps // 2):
            f.write(f"1_{i + 1} 1_{i + 1} 0\n")
    with open(root + ".map", "w") as f:
        for s in range(L):
            f.write(f"{int(bp[s])}\t1.0\t{cm[s]:.10f}\n")
    table = np.array([b"0", b"1"], dtype="S1")
    with gzip.open(root + ".hap.gz", "wb", compresslevel=1) as f:
        cols = haps.T  # [L][n_haps]
        for s in range(L):
            f.write(f"{chrom}:{int(bp[s])}_1_2 SNP_{int(bp[s])} {int(bp[s])} 1 2 ".encode())
            f.write(b" ".join(table[cols[s]].tolist()))
            f.write(b"\n")
    return root


def dataset(root, n_haps, n_sites, span_bp, chrom, seed, founders=None):
    maf = ukbb_maf(chrom, n_sites)
    bp, cm = make_sites(n_sites, span_bp)
    haps = make_haplotypes(n_haps, maf, cm, seed, n_founders=founders)
    write_dataset(root, haps, bp, cm, chrom)
    return haps, bp, cm
