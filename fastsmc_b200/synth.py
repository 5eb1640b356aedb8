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


# ---- large data sets without gz text: the packed-matrix cache written directly (Data::loadBitCache's format) ------------

def _packed_block(args):
    """Worker: haplotypes [lo, hi) of the mosaic model, bit-packed ([n][words] uint64, RAW alleles) + per-site allele counts."""
    lo, hi, maf, p_switch, founders_seed, n_founders, seed, flip = args
    L = len(maf)
    frng = np.random.Generator(np.random.PCG64(founders_seed))
    founders = (frng.random((n_founders, L), dtype=np.float32) < maf[None, :].astype(np.float32)).astype(np.uint8)
    rng = np.random.Generator(np.random.PCG64([seed, lo]))
    n = hi - lo
    sw = rng.random((n, L), dtype=np.float32) < p_switch[None, :].astype(np.float32)
    seg = np.cumsum(sw, axis=1, dtype=np.int32) - 1
    choice = rng.integers(0, n_founders, size=(n, int(seg.max()) + 1), dtype=np.int32)
    h = founders[np.take_along_axis(choice, seg, axis=1), np.arange(L)[None, :]]
    n_flip = rng.binomial(n * L, flip)
    if n_flip:
        at = rng.integers(0, n * L, size=n_flip)
        h.reshape(-1)[at] ^= 1
    words = (L + 63) // 64
    padded = np.zeros((n, words * 64), np.uint8)
    padded[:, :L] = h
    packed = np.packbits(padded.reshape(n, words, 64), axis=2, bitorder="little").view(np.uint64).reshape(n, words)
    return lo, packed, h.sum(axis=0, dtype=np.int64)


def packed_dataset(root, n_haps, n_sites, span_bp, chrom, seed, founders=None, fold=True, csfs=True, processes=None, block=8192):
    """A data set of BASELINE.json's large shapes, written as <root>.samples, <root>.map, a PLACEHOLDER <root>.hap.gz and the
    packed-matrix cache <root>.hap.gz.fsmcbits that Data loads with DecodingParams::hapBitCache (the gz text of 487 409
    samples x 11 528 SNPs would be 11 GB).  Same mosaic model as make_haplotypes; blocks are generated on `processes` host
    processes.  Returns the raw packed matrix [n_haps][words] uint64."""
    import multiprocessing as mp
    import struct
    maf = ukbb_maf(chrom, n_sites)
    bp, cm = make_sites(n_sites, span_bp)
    F = founders or max(64, n_haps // 50)
    dcm = np.diff(cm, prepend=cm[0])
    p_switch = 1.0 - np.exp(-dcm / 2.0)
    p_switch[0] = 1.0
    jobs = [(lo, min(n_haps, lo + block), maf, p_switch, seed, F, seed + 1, 1e-3) for lo in range(0, n_haps, block)]
    words = (n_sites + 63) // 64
    raw = np.zeros((n_haps, words), np.uint64)
    derived = np.zeros(n_sites, np.int64)
    with mp.get_context("fork").Pool(processes or os.cpu_count()) as pool:
        for lo, packed, counts in pool.imap_unordered(_packed_block, jobs):
            raw[lo:lo + len(packed)] = packed
            derived += counts
    os.makedirs(os.path.dirname(os.path.abspath(root)), exist_ok=True)
    with open(root + ".samples", "w") as f:
        f.write("ID_1 ID_2 missing\n0 0 0\n")
        f.write("".join(f"1_{i + 1} 1_{i + 1} 0\n" for i in range(n_haps // 2)))
    with open(root + ".map", "w") as f:
        f.write("".join(f"{int(bp[s])}\t1.0\t{cm[s]:.10f}\n" for s in range(n_sites)))
    with gzip.open(root + ".hap.gz", "wb") as f:
        f.write(b"placeholder: the alleles are in the packed-matrix cache next to this file\n")
    # ---- folding, counts and positions exactly as Data::addSite / addMarker compute them ------------------------------
    total = n_haps
    minor_is_one = (derived <= total - derived) if fold else np.ones(n_sites, bool)
    flip_bits = np.zeros(words * 64, np.uint8)
    flip_bits[:n_sites] = ~minor_is_one
    flip_mask = np.packbits(flip_bits.reshape(words, 64), axis=1, bitorder="little").view(np.uint64).reshape(words)
    folded = raw ^ flip_mask[None, :]
    if n_sites % 64:  # bits beyond the last site stay clear
        folded[:, -1] &= np.uint64((1 << (n_sites % 64)) - 1)
    derived_counts = (np.minimum(derived, total - derived) if fold else derived).astype(np.int32)
    gen = (cm / 100.0).astype(np.float32)
    gd = np.diff(gen.astype(np.float64))
    rate = (gd / np.diff(bp)).astype(np.float32)
    rec = np.concatenate([rate[:1], rate]) if n_sites > 1 else np.zeros(0, np.float32)
    key = []
    for path in (root + ".hap.gz", root + ".samples", root + ".map"):
        st = os.stat(path)
        key.append((st.st_size, int(st.st_mtime)))
    with open(root + ".hap.gz.fsmcbits", "wb") as f:
        f.write(b"FSMCBIT2")
        f.write(struct.pack("<6q2i", key[0][0], key[1][0], key[2][0], key[0][1], key[1][1], key[2][1], int(fold), int(csfs)))
        f.write(struct.pack("<4q", n_haps // 2, n_sites, chrom, words))
        for arr in (np.ascontiguousarray(folded).reshape(-1), flip_mask, np.full(n_sites, total, np.int32), derived_counts,
                    bp.astype(np.int32), gen, rec):
            f.write(struct.pack("<Q", arr.size))
            arr.tofile(f)
    return raw
