"""End-to-end GPU parity through the reference-facing interface (pyASMC: DecodingParams -> FastSMC.run / ASMC.decodePairs)
against the CPU oracle on the reference's example data set, in the reference's own regression configurations
(ASMC_SRC/TESTS/test_fastsmc_regression.cpp:32-160)."""
import gzip

import numpy as np
import pytest

from conftest import ASMC_EXAMPLE, DQ_69, FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, GOLDEN, REGRESSION_PARAMS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def asmc():
    from fastsmc_b200 import asmc as mod
    return mod


def _params(asmc, out, **kw):
    p = asmc.DecodingParams()
    p.verbose = False
    p.inFileRoot = FASTSMC_EXAMPLE
    p.decodingQuantFile = FASTSMC_EXAMPLE_DQ
    p.outFileRoot = out
    p.decodingModeString = "array"
    p.foldData = True
    p.usingCSFS = True
    p.FastSMC = True
    p.hashing = True
    for k, v in dict(REGRESSION_PARAMS, **kw).items():
        setattr(p, k, v)
    p.validateParamsFastSMC()
    return p


def _lines(path):
    with gzip.open(path, "rt") as f:
        return f.read().splitlines()


def _diff(got, want):
    """Empty string when the files agree; otherwise a short description of how they differ."""
    if got == want:
        return ""
    sg, sw = set(got), set(want)
    msg = [f"{len(got)} lines vs {len(want)}; same multiset: {sorted(got) == sorted(want)}; only in got {len(sg - sw)}, "
           f"only in want {len(sw - sg)}"]
    msg += ["+ " + x for x in sorted(sg - sw)[:5]] + ["- " + x for x in sorted(sw - sg)[:5]]
    for i, (a, b) in enumerate(zip(got, want)):
        if a != b:
            msg += [f"first difference at line {i}:", "got  " + a, "want " + b]
            break
    return "\n".join(msg)


CASES = {"hashing": dict(hashing=True), "no_hashing_job7of9": dict(hashing=False, jobInd=7, jobs=9)}


@pytest.mark.parametrize("case", list(CASES))
def test_fastsmc_run_exact_mode_output_identical_to_oracle(asmc, oracle_mod, tmp_path, case):
    """FastSMC.run() with exactArithmetic writes the same .ibd.gz text, line for line, as the oracle's NO_SSE flavour:
    GPU seeding + host order replay + GPU decode + host formatting == the reference pipeline."""
    kw = CASES[case]
    p = _params(asmc, str(tmp_path / "gpu"), exactArithmetic=True, **kw)
    f = asmc.FastSMC(p)
    f.run()
    got = _lines(f"{p.outFileRoot}.{p.jobInd}.{p.jobs}.FastSMC.ibd.gz")
    o = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, str(tmp_path / "o"), **kw, **REGRESSION_PARAMS)
    ref_path = str(tmp_path / "oracle.ibd.gz")
    n = o.run(ref_path)
    want = _lines(ref_path)
    assert len(got) == len(want) == n
    assert got == want


def test_fastsmc_run_fast_mode_within_tolerance(asmc, oracle_mod, tmp_path):
    """Default (FMA) arithmetic: identical segment boundaries on this data set, score / posterior mean within 1e-4."""
    p = _params(asmc, str(tmp_path / "gpu"))
    f = asmc.FastSMC(p)
    f.run()
    got = [l.split("\t") for l in _lines(f"{p.outFileRoot}.1.1.FastSMC.ibd.gz")]
    o = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, str(tmp_path / "o"), hashing=True, **REGRESSION_PARAMS)
    ref_path = str(tmp_path / "oracle.ibd.gz")
    o.run(ref_path)
    want = [l.split("\t") for l in _lines(ref_path)]
    assert [g[:9] for g in got] == [w[:9] for w in want]
    g = np.array([[float(x) for x in r[9:12]] for r in got])
    w = np.array([[float(x) for x in r[9:12]] for r in want])
    np.testing.assert_allclose(g, w, rtol=1e-4)


def test_seeding_candidate_set_bit_exact(asmc, oracle_mod, tmp_path):
    """GPU seeding: the candidate multiset (hapA, hapB, from, to) equals the oracle's; in reference-order mode the
    sequence is identical too; in canonical mode it is sorted by (end word, hapA, hapB)."""
    o = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, str(tmp_path / "o"), hashing=True, **REGRESSION_PARAMS)
    want = o.seed().astype(np.int64)
    for reference_order in (True, False):
        p = _params(asmc, str(tmp_path / f"gpu{int(reference_order)}"), referenceCandidateOrder=reference_order)
        f = asmc.FastSMC(p)
        f.setKeepCandidates(True)
        f.run()
        got = f.getCandidates().astype(np.int64)
        st = f.getSeedingStats()
        assert st.candidates == len(want) == len(got)
        assert st.device.numWords == o.sites // 64 and st.device.pairVisits > st.device.numStarts > 0
        if reference_order:
            assert np.array_equal(got, want)
        else:
            key = lambda a: np.lexsort((a[:, 1], a[:, 0], a[:, 3]))
            assert np.array_equal(got, want[key(want)])
            assert np.array_equal(got, got[key(got)])


def test_binary_output_round_trip(asmc, tmp_path):
    """--bin output read back with BinaryDataReader gives the text output's records (ref: HMM.cpp:1147-1175)."""
    p = _params(asmc, str(tmp_path / "txt"), exactArithmetic=True)
    asmc.FastSMC(p).run()
    text = _lines(f"{p.outFileRoot}.1.1.FastSMC.ibd.gz")
    q = _params(asmc, str(tmp_path / "bin"), exactArithmetic=True, BIN_OUT=True)
    asmc.FastSMC(q).run()
    r = asmc.BinaryDataReader(f"{q.outFileRoot}.1.1.FastSMC.bibd.gz")
    out = []
    while r.moreLinesInFile():
        out.append(r.getNextLine().toString())
    assert len(out) == len(text), (len(out), len(text))
    assert [l.split("\t")[:9] for l in out] == [l.split("\t")[:9] for l in text]
    a = np.array([[float(x) for x in l.split("\t")[9:]] for l in out])
    b = np.array([[float(x) for x in l.split("\t")[9:]] for l in text])
    np.testing.assert_allclose(a, b, rtol=1e-6)  # the binary score is narrowed to float before printing at 7 digits


def test_asmc_decode_pairs_per_site_outputs(asmc, oracle_mod):
    """ASMC.decodePairs (ASMC_SRC/TESTS/test_ASMC.cpp:45-66 shape): per-site posterior mean and MAP of chosen
    haplotype pairs vs the oracle, bit-exact in exact mode, 1e-4 otherwise."""
    a_idx, b_idx = [1, 2, 3, 17, 40], [2, 3, 4, 90, 41]
    for exact in (True, False):
        p = asmc.DecodingParams(ASMC_EXAMPLE, DQ_69, "/tmp/fsmc_asmc_test", 1, 1, "array", False, True, False, False, 0.0,
                                False, True, False, "", False, True)
        p.useKnownSeed = True
        p.exactArithmetic = exact
        p.verbose = False
        m = asmc.ASMC(p)
        m.decodePairs(a_idx, b_idx, False, False, True, True)
        res = m.get_ref_of_results()
        o = oracle_mod.Oracle(ASMC_EXAMPLE, DQ_69, "/tmp/x", hashing=False, FastSMC=False, asmcMode=True, batchSize=64,
                              useKnownSeed=True)
        mean, mp, _ = o.decode_summary(np.array(a_idx), np.array(b_idx))
        got_mean, got_map = np.array(res.per_pair_posterior_means), np.array(res.per_pair_MAPs)
        assert got_mean.shape == mean.shape == (5, o.sites)
        assert [t[0] for t in res.per_pair_indices] == a_idx and [t[2] for t in res.per_pair_indices] == b_idx
        assert res.per_pair_indices[0][1].endswith("#2") and res.per_pair_indices[0][3].endswith("#1")
        if exact:
            assert np.array_equal(got_mean.view(np.uint32), mean.view(np.uint32))
            assert np.array_equal(got_map, mp)
        else:
            np.testing.assert_allclose(got_mean, mean, rtol=1e-4)
            assert (got_map != mp).mean() < 1e-3
        assert np.array_equal(np.array(res.min_posterior_means), got_mean.min(axis=0))
        assert np.array_equal(np.array(res.argmin_posterior_means), got_mean.argmin(axis=0))


def test_all_jobs_partitioned_over_host_threads(asmc, oracle_mod, tmp_path):
    """The jobs/jobInd partition (ref: Data.cpp:62-80, FastSMC_example_multiple_jobs.sh): J=4 jobs of the example data
    dealt to two host threads (two contexts; on a multi-GPU box: two GPUs).  Every job's file equals the oracle's for that
    jobInd, and the union of the jobs' pairs is the single-job run's."""
    p = _params(asmc, str(tmp_path / "gpu"), exactArithmetic=True, jobs=4)
    reports = asmc.pyASMC.runAllJobs(p, [0, 0])
    assert [r.jobInd for r in reports] == [1, 2, 3, 4] and all(r.error == "" for r in reports)
    total = 0
    for r in reports:
        got = _lines(f"{p.outFileRoot}.{r.jobInd}.4.FastSMC.ibd.gz")
        o = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, str(tmp_path / "o"), hashing=True, jobs=4,
                              jobInd=r.jobInd, **REGRESSION_PARAMS)
        ref_path = str(tmp_path / f"oracle{r.jobInd}.ibd.gz")
        n = o.run(ref_path)
        assert _diff(got, _lines(ref_path)) == "", f"job {r.jobInd}"
        assert len(got) == n == r.segments
        assert r.candidates == len(o.candidates())
        total += r.candidates
    one = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, str(tmp_path / "o1"), hashing=True, **REGRESSION_PARAMS)
    assert total == len(one.seed())  # the four triangles tile the pair matrix exactly


def test_asmc_decode_pairs_full_posteriors_and_sum(asmc, oracle_mod):
    """ASMC.decodePairs(per_pair_posteriors=True, sum_of_posteriors=True): per pair a states x sites matrix of
    posterior * expectedCoalTimes[k], and its sum over the pairs (ref: HMM.cpp:1372-1389, TESTS/test_ASMC.cpp:45-66)."""
    a_idx, b_idx = [1, 2, 3, 17, 40, 7], [2, 3, 4, 90, 41, 7 + 12]
    p = asmc.DecodingParams(ASMC_EXAMPLE, DQ_69, "/tmp/fsmc_asmc_test", 1, 1, "array", False, True, False, False, 0.0,
                            False, True, False, "", False, True)
    p.useKnownSeed = True
    p.exactArithmetic = True
    p.verbose = False
    m = asmc.ASMC(p)
    m.decodePairs(a_idx, b_idx, True, True, True, True)
    res = m.get_ref_of_results()
    o = oracle_mod.Oracle(ASMC_EXAMPLE, DQ_69, "/tmp/x", hashing=False, FastSMC=False, asmcMode=True, batchSize=64,
                          useKnownSeed=True)
    post = o.decode_posterior(np.array(a_idx), np.array(b_idx))  # [pair][site][state]
    times = o.vector("expectedTimes").astype(np.float32)
    want = (post * times[None, None, :]).transpose(0, 2, 1)  # float32 product, as the reference stores it
    got = [np.array(x) for x in res.per_pair_posteriors]
    assert len(got) == len(a_idx) and got[0].shape == (o.states, o.sites)
    for g, w in zip(got, want):
        assert np.array_equal(g.view(np.uint32), np.ascontiguousarray(w).view(np.uint32))
    total = np.array(res.sum_of_posteriors)
    assert total.shape == (o.states, o.sites)
    np.testing.assert_allclose(total, want.astype(np.float64).sum(axis=0), rtol=1e-5)
    # the per-site posterior mean is the column sum of a pair's matrix
    np.testing.assert_allclose(np.array(res.per_pair_posterior_means), want.astype(np.float64).sum(axis=1), rtol=1e-5)


def test_asmc_decode_all_posterior_sums_vs_reference_golden(asmc):
    """ASMC decodeAll with doPosteriorSums on the bundled 150-sample array data (all 44 850 haplotype pairs) against
    the reference's own golden sumOverPairs (ASMC_SRC/TESTS/test_regression.cpp:23-67, golden G4, printed at 6
    significant digits).  The CPU oracle reproduces the golden to 7.8e-6 (L1) / 1.2e-4 (worst element)
    (tests/probes/g4_oracle_check.py); the GPU path (FMA arithmetic, float atomics) is held to the north-star 1e-4 on
    the L1 difference.  Row sums are exact properties: every site's posteriors of all pairs sum to the number of
    pairs.  The major/minor sums partition the total."""
    import gzip
    import os
    G4_TOL = 1e-4
    p = asmc.DecodingParams(ASMC_EXAMPLE, DQ_69, "", 1, 1, "array", False, True, False, False, 0.0, False, True)
    p.useKnownSeed = True
    p.verbose = False
    p.doMajorMinorPosteriorSums = True
    r = asmc.ASMC(p).decodeAllInJob()
    got = np.array(r.sumOverPairs, dtype=np.float64)
    gold = np.loadtxt(gzip.open(os.path.join(GOLDEN, "asmc_sum_over_pairs.gz"), "rt"))
    assert got.shape == gold.shape == (6760, 69)
    np.testing.assert_allclose(got.sum(axis=1), 44850.0, rtol=2e-4)
    l1 = np.abs(got - gold).sum() / gold.sum()
    worst = (np.abs(got - gold) / np.maximum(gold, 1e-3)).max()
    print(f"G4: L1 relative difference {l1:.3e}, worst element {worst:.3e}")
    assert l1 < G4_TOL and worst < 2e-3, (l1, worst)
    parts = np.array(r.sumOverPairs00, dtype=np.float64) + np.array(r.sumOverPairs01) + np.array(r.sumOverPairs11)
    np.testing.assert_allclose(parts, got, rtol=1e-5, atol=1e-4)
    assert np.array(r.sumOverPairs01).sum() > 0 and np.array(r.sumOverPairs11).sum() > 0


def test_hmm_decode_returns_full_posterior_of_one_pair(asmc, oracle_mod):
    """HMM::decode(obs, from, to) (ref: HMM.cpp:1464-1495): states x sites posterior of one pair, bit-identical to the
    oracle in exact mode; decodeSummarize's posterior mean is its expectation."""
    p = asmc.DecodingParams(ASMC_EXAMPLE, DQ_69, "", 1, 1, "array", False, True)
    p.useKnownSeed = True
    p.exactArithmetic = True
    p.verbose = False
    hmm = asmc.HMM(asmc.Data(p), p)
    obs = hmm.makePairObs(1, 0, 2, 5)  # haplotype 1 of individual 0 vs haplotype 2 of individual 5
    got = np.array(hmm.decode(obs, 100, 900), dtype=np.float32)
    o = oracle_mod.Oracle(ASMC_EXAMPLE, DQ_69, "/tmp/x", hashing=False, FastSMC=False, asmcMode=True, batchSize=64,
                          useKnownSeed=True)
    want = o.decode_posterior(np.array([0]), np.array([11]), 100, 900)[0].T
    assert got.shape == want.shape == (o.states, 800)
    assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(want).view(np.uint32))
    np.testing.assert_allclose(got.sum(axis=0), 1.0, rtol=1e-5)
    whole = np.array(hmm.decode(obs), dtype=np.float32)
    mean, _ = hmm.decodeSummarize(obs)
    np.testing.assert_allclose((whole * o.vector("expectedTimes")[:, None]).sum(axis=0), np.array(mean), rtol=1e-5)


def test_pipelined_decode_workers_give_the_same_file(asmc, tmp_path, monkeypatch):
    """Chunks are decoded by two workers with their own contexts when the narrow kernel runs (FastSMC_exe's default
    conditional age estimates); they complete in submission order, so the output is byte-identical to the one-worker
    run.  Small chunks (two reference batches each) so that the example data makes many of them."""
    monkeypatch.setenv("FSMC_FLUSH_BATCHES", "2")
    texts = {}
    for workers in ("1", "2"):
        monkeypatch.setenv("FSMC_DECODE_WORKERS", workers)
        for hashing in (True, False):
            kw = dict(hashing=True) if hashing else dict(hashing=False, jobInd=7, jobs=9)
            p = _params(asmc, str(tmp_path / f"w{workers}h{int(hashing)}"), noConditionalAgeEstimates=False, **kw)
            f = asmc.FastSMC(p)
            f.run()
            st = f.hmm().getRunStats()
            assert st.decodeCalls >= 8 and st.segments == f.hmm().getNumberOfDetectedSegments() > 0
            texts[workers, hashing] = _lines(f"{p.outFileRoot}.{p.jobInd}.{p.jobs}.FastSMC.ibd.gz")
    for hashing in (True, False):
        assert len(texts["1", hashing]) > 100 and texts["1", hashing] == texts["2", hashing]


# ---- reference candidate order computed on the device (csrc/seed_order.cu), synthetic data ------------------------------


def _synthetic_params(asmc, root, out, dq, **kw):
    p = asmc.DecodingParams()
    p.verbose = False
    p.inFileRoot, p.decodingQuantFile, p.outFileRoot = root, dq, out
    p.decodingModeString, p.foldData, p.usingCSFS, p.FastSMC, p.hashing = "array", True, True, True, True
    for k, v in dict(REGRESSION_PARAMS, **kw).items():
        setattr(p, k, v)
    p.validateParamsFastSMC()
    return p


@pytest.mark.parametrize("n_haps,n_sites,gap,min_m,two_pass", [(240, 3200, 1, 0.0, False), (400, 1920, 0, 0.05, True),
                                                              (150, 6400, 2, 0.4, False)])
def test_device_candidate_order_equals_literal_replay_on_dense_matches(asmc, tmp_path, monkeypatch, n_haps, n_sites, gap,
                                                                       min_m, two_pass):
    """Few founders: tens of thousands of short intervals, most of them never candidates but all of them nodes of the
    reference's extend map; the map grows through many rehashes, buckets empty and refill.  The device ordering inside
    fsmc_seed must hand the candidates over in the order of the LITERAL replay of the two boost maps (linked node lists,
    CandidateOrder.hpp::replayReferenceOrder) over brute-force intervals."""
    from fastsmc_b200 import synth
    from test_host_layer import _brute_force_intervals
    if two_pass:
        monkeypatch.setenv("FSMC_ORDER_TWO_PASS", "1")  # creation sort as two stable passes (keys wider than 64 bits)
    root = str(tmp_path / "dense")
    synth.dataset(root, n_haps, n_sites, 3000 * n_sites, 1, 77 + n_haps, founders=6)
    p = _synthetic_params(asmc, root, root + ".out", FASTSMC_EXAMPLE_DQ, gap=gap, min_m=min_m)
    d = asmc.Data(p)
    iv = _brute_force_intervals(np.array(d.hapBits), d.sites // 64, gap)
    literal = np.array(asmc.pyASMC.replayReferenceOrder(iv, d, gap, min_m, fast=False))
    want = iv[literal]
    want = np.stack([want[:, 0], want[:, 1], want[:, 2] * 64, want[:, 3] * 64 + 63], axis=1)
    f = asmc.FastSMC(p)
    f.setKeepCandidates(True)
    f.run()
    got = f.getCandidates().astype(np.int64)
    st = f.getSeedingStats().device
    assert st.numIntervals == len(iv) and st.orderEpochs > 5 and len(want) > 100
    assert len(got) == len(want) == st.numMatches
    assert np.array_equal(got, want)


def test_hashing_jobs_on_synthetic_data_equal_oracle(asmc, oracle_mod, tmp_path):
    """cfg3-style hashing run on synthetic data, cut into jobs (jobs/jobInd windows, the last job with the remainder):
    the decodeFromHashing call stream equals the oracle's (which walks real linked lists), and in exact mode every job's
    .ibd.gz equals the oracle's line for line."""
    from fastsmc_b200 import synth
    root = str(tmp_path / "syn")
    synth.dataset(root, 1300, 6400, 64_000_000, 1, 4242)
    for jobs, job_ind in ((4, 2), (4, 4), (1, 1)):
        o = oracle_mod.Oracle(root, DQ_69, str(tmp_path / "o"), hashing=True, jobs=jobs, jobInd=job_ind, **REGRESSION_PARAMS)
        want = o.seed().astype(np.int64)
        p = _synthetic_params(asmc, root, str(tmp_path / f"gpu{jobs}_{job_ind}"), DQ_69, jobs=jobs, jobInd=job_ind,
                              exactArithmetic=True)
        f = asmc.FastSMC(p)
        f.setKeepCandidates(True)
        f.run()
        got = f.getCandidates().astype(np.int64)
        assert len(want) > 200 and np.array_equal(got, want)
        ref_path = str(tmp_path / f"oracle{jobs}_{job_ind}.ibd.gz")
        n = o.run(ref_path)
        mine = _lines(f"{p.outFileRoot}.{job_ind}.{jobs}.FastSMC.ibd.gz")
        assert len(mine) == n and mine == _lines(ref_path)


@pytest.mark.parametrize("max_seeds", [0, 30], ids=["plain", "max_seeds30"])
def test_device_candidate_order_at_cfg3_density(asmc, oracle_mod, tmp_path, max_seeds):
    """2 000 diploid samples x 50 000 SNPs at UKBB chr1 array density (cfg3's shape, a fifth of its samples): millions of
    intervals, the extend map rehashes up to millions of buckets.  Candidate stream vs the oracle's seeding; also with
    sub-hashing of the buckets above 30 haplotypes (max_seeds), whose registrations run ahead of the current word."""
    from fastsmc_b200 import synth
    root = str(tmp_path / "cfg3s")
    synth.dataset(root, 4000, 50_000, 240_000_000, 1, 20201117 + 3)
    o = oracle_mod.Oracle(root, DQ_69, str(tmp_path / "o"), hashing=True, **dict(REGRESSION_PARAMS, max_seeds=max_seeds))
    want = o.seed().astype(np.int64)
    p = _synthetic_params(asmc, root, str(tmp_path / "gpu"), DQ_69, max_seeds=max_seeds)
    f = asmc.FastSMC(p)
    f.setKeepCandidates(True)
    f.run()
    got = f.getCandidates().astype(np.int64)
    st = f.getSeedingStats().device
    assert st.numIntervals > 1_000_000 and st.orderEpochs > (15 if max_seeds == 0 else 10)
    assert len(got) == len(want) > 50_000
    assert np.array_equal(got, want)


@pytest.mark.parametrize("options", [dict(skip=0.13), dict(skip=0.145, gap=2, min_m=2.5), dict(skip=0.15, gap=0, min_m=1.0)],
                         ids=["skip0.13", "skip0.145-gap2", "skip0.15-gap0"])
@pytest.mark.parametrize("reference_order", [True, False], ids=["reference-order", "canonical-order"])
def test_low_complexity_word_skipping_equals_oracle(asmc, oracle_mod, tmp_path, options, reference_order):
    """DecodingParams::skip (FastSMC.cpp:208-219, ExtendHash.hpp:102-106): words with few distinct haplotype words seed no
    pairs and extend every live interval.  Dense synthetic data (6 founders): a quarter to a half of the words are skipped.
    Candidate stream bit-exact against the oracle (which equals the reference build on this data,
    tests/test_reference_build.py); in exact mode and reference order the .ibd.gz is identical too."""
    from fastsmc_b200 import synth
    root = str(tmp_path / "dense")
    synth.dataset(root, 200, 1920, 3000 * 1920, 1, 11, founders=6)
    options = dict(dict(min_m=2.0), **options)
    o = oracle_mod.Oracle(root, FASTSMC_EXAMPLE_DQ, str(tmp_path / "o"), hashing=True, **dict(REGRESSION_PARAMS, **options))
    want = o.seed().astype(np.int64)
    p = _synthetic_params(asmc, root, str(tmp_path / "gpu"), FASTSMC_EXAMPLE_DQ, exactArithmetic=True,
                          referenceCandidateOrder=reference_order, **options)
    f = asmc.FastSMC(p)
    f.setKeepCandidates(True)
    f.run()
    got = f.getCandidates().astype(np.int64)
    assert len(want) > 300 and len(got) == len(want)
    if reference_order:
        assert np.array_equal(got, want)
        ref_path = str(tmp_path / "oracle.ibd.gz")
        n = o.run(ref_path)
        mine = _lines(f"{p.outFileRoot}.1.1.FastSMC.ibd.gz")
        assert len(mine) == n and mine == _lines(ref_path)
    else:
        key = lambda a: np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))
        assert np.array_equal(got[key(got)], want[key(want)])  # the same candidate multiset


@pytest.mark.parametrize("options", [dict(max_seeds=20), dict(max_seeds=50, gap=2, min_m=2.5), dict(max_seeds=8, gap=0, min_m=1.0),
                                     dict(max_seeds=8, skip=0.13), dict(max_seeds=3)],
                         ids=["max_seeds20", "max_seeds50-gap2", "max_seeds8-gap0", "max_seeds8-skip0.13", "max_seeds3"])
@pytest.mark.parametrize("reference_order", [True, False], ids=["reference-order", "canonical-order"])
def test_sub_hashing_of_oversized_buckets_equals_oracle(asmc, oracle_mod, tmp_path, options, reference_order):
    """DecodingParams::max_seeds (SeedHash.hpp:56-69, 85-93): a bucket with more than max_seeds haplotypes is re-hashed on
    the following words (inside the read-ahead buffer) and its pairs are extended to the deepest word.  Dense synthetic
    data (6 founders: buckets of ~60 haplotypes), so nearly every word sub-hashes, several levels deep.  Candidate stream
    bit-exact against the oracle (which equals the reference build with these options, tests/test_reference_build.py);
    in exact mode and reference order the .ibd.gz is identical too."""
    from fastsmc_b200 import synth
    root = str(tmp_path / "dense")
    synth.dataset(root, 200, 1920, 3000 * 1920, 1, 11, founders=6)
    options = dict(dict(min_m=2.0), **options)
    o = oracle_mod.Oracle(root, FASTSMC_EXAMPLE_DQ, str(tmp_path / "o"), hashing=True, **dict(REGRESSION_PARAMS, **options))
    want = o.seed().astype(np.int64)
    plain = oracle_mod.Oracle(root, FASTSMC_EXAMPLE_DQ, str(tmp_path / "o0"), hashing=True,
                              **dict(REGRESSION_PARAMS, **dict(options, max_seeds=0))).seed().astype(np.int64)
    if options["max_seeds"] <= 20:
        assert len(want) != len(plain) or not np.array_equal(want, plain)  # the option changes the candidates on this data
    p = _synthetic_params(asmc, root, str(tmp_path / "gpu"), FASTSMC_EXAMPLE_DQ, exactArithmetic=True,
                          referenceCandidateOrder=reference_order, **options)
    f = asmc.FastSMC(p)
    f.setKeepCandidates(True)
    f.run()
    got = f.getCandidates().astype(np.int64)
    assert len(want) > 300 and len(got) == len(want)
    key = lambda a: np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))
    assert np.array_equal(got[key(got)], want[key(want)])  # the same candidate multiset
    if reference_order:
        assert np.array_equal(got, want)
        ref_path = str(tmp_path / "oracle.ibd.gz")
        n = o.run(ref_path)
        mine = _lines(f"{p.outFileRoot}.1.1.FastSMC.ibd.gz")
        assert len(mine) == n and mine == _lines(ref_path)


@pytest.mark.parametrize("max_seeds", [50, 500])
def test_sub_hashing_at_larger_buckets(asmc, oracle_mod, tmp_path, max_seeds):
    """max_seeds in {50, 500} on 4 000 haplotypes drawn from 6 founders (buckets of several hundred haplotypes, i.e. above
    both cut-offs): the candidate stream in reference order equals the oracle's."""
    from fastsmc_b200 import synth
    root = str(tmp_path / "dense")
    synth.dataset(root, 2000, 1280, 3000 * 1280, 1, 5, founders=6)
    options = dict(min_m=2.0, max_seeds=max_seeds)
    o = oracle_mod.Oracle(root, FASTSMC_EXAMPLE_DQ, str(tmp_path / "o"), hashing=True, **dict(REGRESSION_PARAMS, **options))
    want = o.seed().astype(np.int64)
    p = _synthetic_params(asmc, root, str(tmp_path / "gpu"), FASTSMC_EXAMPLE_DQ, **options)
    f = asmc.FastSMC(p)
    f.setKeepCandidates(True)
    f.run()
    got = f.getCandidates().astype(np.int64)
    assert len(want) > 1000 and len(got) == len(want)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("jobs,job_ind", [(4, 2), (4, 3), (4, 4)])
def test_sub_hashing_inside_jobs(asmc, oracle_mod, tmp_path, jobs, job_ind):
    """max_seeds together with the jobs/jobInd partition (Data.cpp:62-80): the buckets are those of the job's haplotypes,
    the job filter applies to the pairs of the final nested buckets.  Candidate stream in reference order vs the oracle."""
    from fastsmc_b200 import synth
    root = str(tmp_path / "dense")
    synth.dataset(root, 300, 1920, 3000 * 1920, 1, 23, founders=6)
    options = dict(min_m=2.0, max_seeds=10, jobs=jobs, jobInd=job_ind)
    o = oracle_mod.Oracle(root, FASTSMC_EXAMPLE_DQ, str(tmp_path / "o"), hashing=True, **dict(REGRESSION_PARAMS, **options))
    want = o.seed().astype(np.int64)
    p = _synthetic_params(asmc, root, str(tmp_path / "gpu"), FASTSMC_EXAMPLE_DQ, **options)
    f = asmc.FastSMC(p)
    f.setKeepCandidates(True)
    f.run()
    got = f.getCandidates().astype(np.int64)
    assert len(want) > 100 and len(got) == len(want)
    assert np.array_equal(got, want)
