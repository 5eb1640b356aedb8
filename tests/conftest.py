import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "data")
FASTSMC_EXAMPLE = os.path.join(GOLDEN, "fastsmc_example", "example")
FASTSMC_EXAMPLE_DQ = os.path.join(GOLDEN, "fastsmc_example", "example.decodingQuantities.gz")
ASMC_EXAMPLE = os.path.join(GOLDEN, "asmc_example", "exampleFile.n300.array")
DQ_69 = os.path.join(DATA, "30-100-2000.decodingQuantities.gz")

# parameters of the reference's regression tests (ASMC_SRC/TESTS/test_fastsmc_regression.cpp:34-52, 99-119)
REGRESSION_PARAMS = dict(batchSize=32, min_m=1.5, FastSMC=True, BIN_OUT=False, outputIbdSegmentLength=True, time=50,
                         noConditionalAgeEstimates=True, doPerPairMAP=True, doPerPairPosteriorMean=True,
                         useKnownSeed=True)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def have_gpu():
    try:
        from fastsmc_b200 import _native
        return _native.lib().fsmc_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


def model_from_oracle(o, pyoracle):
    """C-ABI model tables from an oracle instance (the oracle is the checker, so its tables are the inputs)."""
    gen, _ = o.positions()
    g = gen.astype(np.float32)
    keys = [0.0] + [pyoracle.round_morgans(float(g[i] - g[i - 1])) for i in range(1, len(g))]
    uniq = sorted(set(keys[1:]))
    index = {k: i for i, k in enumerate(uniq)}
    rows = np.array([0] + [index[k] for k in keys[1:]], np.int32)
    tabs = np.stack([o.transition(k) for k in uniq])  # [n][4][S]  (D,B,U,RR)
    e1, e0m1, e2m0 = o.emissions()
    return dict(initial_state_prob=o.vector("initialStateProb"), expected_times=o.vector("expectedTimes"),
                column_ratios=o.vector("columnRatios"), emission1=e1, emission0minus1=e0m1, emission2minus0=e2m0,
                D=tabs[:, 0], B=tabs[:, 1], U=tabs[:, 2], RR=tabs[:, 3], distance_row=rows,
                state_threshold=o.state_threshold, age_threshold=o.age_threshold,
                probability_threshold=o.probability_threshold)


def context_from_oracle(o, pyoracle, device=0):
    from fastsmc_b200 import _native
    ctx = _native.Context(device)
    ctx.set_model(**model_from_oracle(o, pyoracle))
    ctx.set_haplotypes(_native.pack_haplotypes(o.haplotypes()), o.sites)
    return ctx


def check_segments_up_to_threshold_ties(o, got, want_keys, pair_haps, rtol, frm=0, to=None, max_pairs=200):
    """The north star's rule for segment calls: identical, except at sites whose posterior lies within `rtol` (relative) of a
    threshold.  `got`: the GPU's segment records; `want_keys`: the oracle's {(pair, posStart, posEnd)}; pair_haps(pair) ->
    (hapA, hapB).  For every pair whose segment lists differ, the oracle's per-site IBD probability is recomputed and the
    GPU's per-site level (from its records) must equal the oracle's at every site that is not within rtol of a threshold.
    Returns the number of pairs that differed."""
    to = o.sites if to is None else to
    got_keys = {(int(s["pair"]), int(s["posStart"]), int(s["posEnd"])) for s in got}
    differing = sorted({k[0] for k in got_keys ^ set(want_keys)})
    assert len(differing) <= max_pairs, f"{len(differing)} pairs differ"
    thr = np.float32(o.probability_threshold) * np.array([1000, 100, 10, 1], np.float32)
    by_pair = {}
    for s in got:
        if int(s["pair"]) in differing:
            by_pair.setdefault(int(s["pair"]), []).append(s)
    for pair in differing:
        a, b = pair_haps(pair)
        _, _, ibd = o.decode_summary(np.array([a], np.uint32), np.array([b], np.uint32), frm, to, mean=False, map_=False)
        ibd = ibd[0]
        lv = np.full(to - frm, -1)
        for i in (3, 2, 1, 0):
            lv[ibd >= thr[i]] = i
        ambiguous = np.zeros(to - frm, bool)
        for t in thr:
            ambiguous |= np.abs(ibd - t) <= rtol * t
        mine = np.full(to - frm, -1)
        for s in by_pair.get(pair, []):
            mine[int(s["posStart"]) - frm:int(s["posEnd"]) + 1 - frm] = int(s["level"])
        bad = np.flatnonzero((mine != lv) & ~ambiguous)
        assert len(bad) == 0, f"pair {pair}: level differs at sites {bad[:5] + frm} where the posterior is not near a threshold"
    return len(differing)
