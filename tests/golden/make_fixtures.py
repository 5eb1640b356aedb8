"""Regenerates the data fixtures under tests/golden/ and data/ from the read-only reference checkout.

Run in the build container (where /root/reference exists):  python tests/golden/make_fixtures.py

What it writes (all are DATA files of the reference, copied or row-filtered; no source code):
  tests/golden/fastsmc_example/example.{hap.gz,samples,map.gz}        FILES/FASTSMC_EXAMPLE (map thinned, see below)
  tests/golden/fastsmc_example/example.decodingQuantities.gz          row-filtered copy (see below)
  tests/golden/fastsmc_example/regression_output{,_no_hashing}.ibd.gz the reference's golden outputs G1 / G2
  tests/golden/asmc_example/exampleFile.n300.array.{hap.gz,samples,map.gz}   FILES/EXAMPLE
  tests/golden/binary_output.bibd.gz                                  ASMC_SRC/TESTS/data (binary reader fixture)
  tests/golden/asmc_sum_over_pairs.gz                                 ASMC_SRC/TESTS/data/regression_test_original.gz: the
                                                                      reference's golden sumOverPairs (6760 sites x 69 states)
                                                                      of ASMC_SRC/TESTS/test_regression.cpp (golden G4)
  data/30-100-2000.decodingQuantities.gz                              row-filtered copy of FILES/DECODING_QUANTITIES
  data/ukbb_maf.npz                                                   MAF column of FILES/UKBB.frq for chr 1 (first 50k), 20, 22

Row filtering of a decoding-quantities file keeps every line verbatim except
  * rows of the RowRatios/Uvectors/Bvectors/Dvectors sections whose distance key is not needed by the
    bundled example data nor lies in [2e-5, 6e-5] Morgans (the gap range of the synthetic benchmarks), and
  * rows of HomozygousEmissions (sequence mode only; out of scope).
The kept rows are byte-identical to the original, so every parsed float is identical.

Thinning of example.map keeps, for every SNP of example.hap.gz, the map rows that the reference's
interpolation (Data.cpp:523-547) reads for it; the interpolated positions are therefore identical.
"""
import gzip
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def needed_keys(gen):
    from oracle.pyoracle import round_morgans
    g = np.asarray(gen, np.float32)
    return {np.float32(round_morgans(float(g[i] - g[i - 1]))) for i in range(1, len(g))}


def filter_dq(src, dst, keys):
    keyed = {"rowratios", "uvectors", "bvectors", "dvectors"}
    section = None
    kept = dropped = 0
    with gzip.open(src, "rt") as fin, gzip.open(dst, "wt", compresslevel=9) as fout:
        for line in fin:
            tok = line.split()
            head = tok[0].lower() if tok else ""
            if head and not (head[0].isdigit() or head[0] in "-.+"):
                section = head
                fout.write(line)
                continue
            if section in keyed and tok:
                k = np.float32(np.longdouble(tok[0]))
                if k in keys or (np.float32(2e-5) <= k <= np.float32(6e-5)):
                    fout.write(line)
                    kept += 1
                else:
                    dropped += 1
                continue
            if section == "homozygousemissions":
                dropped += 1
                continue
            fout.write(line)
    print(f"{os.path.basename(dst)}: kept {kept} keyed rows, dropped {dropped}")


def thin_map(src, hap_gz, dst):
    bps = []
    with gzip.open(hap_gz, "rt") as f:
        for line in f:
            bps.append(int(line.split(None, 3)[2]))
    rows = []
    with open(src) as f:
        for line in f:
            t = line.split()
            try:
                rows.append((int(t[0]), line))
            except (ValueError, IndexError):
                pass
    keep = set()
    g = 0
    for bp in bps:
        while bp > rows[g][0] and g < len(rows) - 1:
            g += 1
        keep.add(g)
        if g > 0:
            keep.add(g - 1)
    keep.add(0)
    keep.add(len(rows) - 1)
    with gzip.open(dst, "wt", compresslevel=9) as f:
        for i in sorted(keep):
            f.write(rows[i][1])
    print(f"{os.path.basename(dst)}: kept {len(keep)} of {len(rows)} map rows")


def main():
    from oracle.pyoracle import Oracle
    fx = os.path.join(HERE, "fastsmc_example")
    ax = os.path.join(HERE, "asmc_example")
    data = os.path.join(ROOT, "data")
    for d in (fx, ax, data):
        os.makedirs(d, exist_ok=True)
    E = REF + "/FILES/FASTSMC_EXAMPLE/"
    for f in ("example.hap.gz", "example.samples", "regression_output.ibd.gz", "regression_output_no_hashing.ibd.gz"):
        shutil.copyfile(E + f, os.path.join(fx, f))
    thin_map(E + "example.map", E + "example.hap.gz", os.path.join(fx, "example.map.gz"))
    A = REF + "/FILES/EXAMPLE/"
    for f in ("exampleFile.n300.array.hap.gz", "exampleFile.n300.array.samples", "exampleFile.n300.array.map.gz"):
        shutil.copyfile(A + f, os.path.join(ax, f))
    shutil.copyfile(REF + "/ASMC_SRC/TESTS/data/binary_output.bibd.gz", os.path.join(HERE, "binary_output.bibd.gz"))
    shutil.copyfile(REF + "/ASMC_SRC/TESTS/data/regression_test_original.gz", os.path.join(HERE, "asmc_sum_over_pairs.gz"))

    o = Oracle(E + "example", E + "example.decodingQuantities.gz", "/tmp/x", hashing=True, time=50)
    keys = needed_keys(o.positions()[0])
    filter_dq(E + "example.decodingQuantities.gz", os.path.join(fx, "example.decodingQuantities.gz"), keys)
    # the thinned map must interpolate to the same positions
    o2 = Oracle(os.path.join(fx, "example"), os.path.join(fx, "example.decodingQuantities.gz"), "/tmp/x", hashing=True,
                time=50)
    assert np.array_equal(o.positions()[0], o2.positions()[0]) and np.array_equal(o.positions()[1], o2.positions()[1])
    assert all(np.array_equal(x, y) for x, y in zip(o.emissions(), o2.emissions()))

    Q = REF + "/FILES/DECODING_QUANTITIES/30-100-2000.decodingQuantities.gz"
    oa = Oracle(A + "exampleFile.n300.array", Q, "/tmp/x", hashing=False, FastSMC=False, asmcMode=True, batchSize=64)
    filter_dq(Q, os.path.join(data, "30-100-2000.decodingQuantities.gz"), needed_keys(oa.positions()[0]))

    maf = {}
    with open(REF + "/FILES/UKBB.frq") as f:
        next(f)
        for line in f:
            t = line.split()
            if t[0] in ("1", "20", "22"):
                maf.setdefault(t[0], []).append(float(t[4]))
    np.savez_compressed(os.path.join(data, "ukbb_maf.npz"), chr1=np.array(maf["1"][:50000], np.float32),
                        chr20=np.array(maf["20"], np.float32), chr22=np.array(maf["22"], np.float32))
    print({k: len(v) for k, v in maf.items()})


if __name__ == "__main__":
    main()
