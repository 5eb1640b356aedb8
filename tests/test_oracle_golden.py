"""Pins the CPU oracle against the reference's own golden outputs and known-answer tests.

G1  FILES/FASTSMC_EXAMPLE/regression_output.ibd.gz             (ASMC_SRC/TESTS/test_fastsmc_regression.cpp:32-94)
G2  FILES/FASTSMC_EXAMPLE/regression_output_no_hashing.ibd.gz  (ASMC_SRC/TESTS/test_fastsmc_regression.cpp:97-160)
K*  ASMC_SRC/TESTS/test_hmm_utils.cpp:180-332, ASMC_SRC/TESTS/test_hashing.cpp:120-151

The goldens were produced by an AVX build (approximate-reciprocal posterior normalisation) linked
against libstdc++ <= 10 (divide-and-reject std::shuffle); the oracle reproduces them line for line
in that flavour (simdFlavor=True, shuffleFlavor=2).  The parity target of the CUDA path is the oracle's NO_SSE
flavour (exact 1/x), which differs from the goldens only within the 4e-4 band of the approximate reciprocal —
also checked here.
"""
import gzip
import os

import numpy as np
import pytest

from conftest import FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, GOLDEN, REGRESSION_PARAMS

CASES = {
    "G1": ("regression_output.ibd.gz", dict(hashing=True), 1524),
    "G2": ("regression_output_no_hashing.ibd.gz", dict(hashing=False, jobInd=7, jobs=9), 2986),
}


def _lines(path):
    with gzip.open(path, "rt") as f:
        return f.read().splitlines()


@pytest.mark.parametrize("case", ["G1", "G2"])
def test_oracle_reproduces_golden_line_for_line(oracle_mod, tmp_path, case):
    name, extra, n_lines = CASES[case]
    o = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, str(tmp_path / "o"), simdFlavor=True, shuffleFlavor=2,
                          **extra, **REGRESSION_PARAMS)
    out = str(tmp_path / "out.ibd.gz")
    assert o.run(out, 0) == n_lines
    want = _lines(os.path.join(GOLDEN, "fastsmc_example", name))
    got = _lines(out)
    assert len(want) == n_lines
    assert got == want


def test_oracle_nosse_flavour_within_rcp_band_of_golden(oracle_mod, tmp_path):
    """Exact-reciprocal arithmetic (the CUDA path's parity target) vs the AVX-built golden G1: same records, the
    float columns within the approximate reciprocal's error band (rel. 1.5 * 2^-12)."""
    name, extra, n_lines = CASES["G1"]
    o = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, str(tmp_path / "o"), simdFlavor=False, shuffleFlavor=2,
                          **extra, **REGRESSION_PARAMS)
    out = str(tmp_path / "out.ibd.gz")
    assert o.run(out, 0) == n_lines
    want = [l.split("\t") for l in _lines(os.path.join(GOLDEN, "fastsmc_example", name))]
    got = [l.split("\t") for l in _lines(out)]
    # ids, haps, chr, bp start/end: identical except where a site's IBD probability sits within the reciprocal's
    # error band of a threshold (a boundary then moves by a site, or a segment splits differently)
    same = [i for i, (w, g) in enumerate(zip(want, got)) if w[:9] == g[:9]]
    assert len(same) >= 0.99 * n_lines
    w = np.array([[float(x) for x in want[i][9:12]] for i in same])
    g = np.array([[float(x) for x in got[i][9:12]] for i in same])
    np.testing.assert_allclose(g, w, rtol=4e-4)


def test_shuffle_flavours(oracle_mod):
    """Flavour 1 (Lemire written out) must equal flavour 0 (this libstdc++'s std::shuffle) on libstdc++ >= 11."""
    kw = dict(hashing=True, **REGRESSION_PARAMS)
    a = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, "/tmp/x", shuffleFlavor=0, **kw).undistinguished()
    b = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, "/tmp/x", shuffleFlavor=1, **kw).undistinguished()
    assert np.array_equal(a, b)


# ---- known-answer tests of the reference's unit tests --------------------------------------------------------


def test_round_morgans_kat(oracle_mod):  # ASMC_SRC/TESTS/test_hmm_utils.cpp:180-202
    rm = oracle_mod.round_morgans
    f = np.float32
    for prec in (0, 1, 2):
        assert rm(0.4, prec, 0.5) == 0.5
    assert rm(0.0, 5, 0.5) == 0.5 and rm(-1.0, 7, 0.5) == 0.5
    a = 0.123456
    assert f(rm(a, 0, 1e-10)) == f(0.1)
    assert f(rm(a, 1, 1e-10)) == f(0.12)
    assert f(rm(a, 2, 1e-10)) == f(0.123)
    assert f(rm(a, 3, 1e-10)) == f(0.1235)
    assert f(rm(a, 4, 1e-10)) == f(0.12346)


def test_round_physical_kat(oracle_mod):  # ASMC_SRC/TESTS/test_hmm_utils.cpp:204-229
    rp = oracle_mod.round_physical
    for v in (-1, 0, 1):
        for prec in (0, 1, 2):
            assert rp(v, prec) == 1
    assert [rp(123456, p) for p in range(6)] == [100000, 120000, 123000, 123500, 123460, 123456]


def test_from_to_position_kat(oracle_mod):  # ASMC_SRC/TESTS/test_hmm_utils.cpp:298-332
    gen = [0.12, 0.23, 0.34, 0.45, 0.56, 0.67]
    gf = oracle_mod.get_from_position
    assert [gf(gen, 4, d) for d in (1, 21, 23, 30, 45, 60)] == [3, 2, 1, 1, 0, 0]
    assert [gf(gen, 0, d) for d in (1e-6, 1, 10)] == [0, 0, 0]
    gt = oracle_mod.get_to_position
    assert [gt(gen, 1, d) for d in (1, 10, 12, 30, 40, 60)] == [3, 3, 4, 5, 6, 6]
    assert [gt(gen, 6, d) for d in (1e-6, 1, 10)] == [6, 6, 6]


def test_cm_between_kat(oracle_mod):  # ASMC_SRC/TESTS/test_hashing.cpp:120-151
    gen = [0.1 * i for i in range(9)]
    cb = oracle_mod.cm_between
    assert cb(0, 0, gen, 4) == pytest.approx(100.0 * (gen[3] - gen[0]), rel=1e-6)
    assert cb(0, 1, gen, 4) == pytest.approx(100.0 * (gen[7] - gen[0]), rel=1e-6)
    assert cb(1, 5, gen, 4) == pytest.approx(100.0 * (gen[8] - gen[4]), rel=1e-6)  # end clamps to the last site


def test_posteriors_are_distributions(oracle_mod):
    """Per-site posteriors are not pinned by any reference test; sanity: each is a distribution and the
    summaries agree with it."""
    o = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, "/tmp/x", hashing=True, **REGRESSION_PARAMS)
    a, b = np.array([0, 5, 17], np.uint32), np.array([3, 9, 200], np.uint32)
    post = o.decode_posterior(a, b, 100, 400)
    np.testing.assert_allclose(post.sum(axis=2), 1.0, rtol=2e-5)
    mean, mp, ibd = o.decode_summary(a, b, 100, 400)
    np.testing.assert_allclose(mean, post @ o.vector("expectedTimes"), rtol=1e-4)
    assert np.array_equal(mp, post.argmax(axis=2))
    np.testing.assert_allclose(ibd, post[:, :, :o.state_threshold].sum(axis=2), rtol=1e-4, atol=1e-12)
