"""Host layer (C++ classes behind pyASMC) against the CPU oracle and the reference's known-answer tests.  No GPU."""
import gzip
import os

import numpy as np
import pytest

from conftest import DQ_69, FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, GOLDEN, REGRESSION_PARAMS


@pytest.fixture(scope="module")
def asmc():
    from fastsmc_b200 import asmc as mod
    return mod


def _params(asmc, **kw):
    p = asmc.DecodingParams()
    p.verbose = False
    p.inFileRoot = FASTSMC_EXAMPLE
    p.decodingQuantFile = FASTSMC_EXAMPLE_DQ
    p.outFileRoot = "/tmp/fsmc_host_test"
    p.decodingModeString = "array"
    p.foldData = True
    p.usingCSFS = True
    p.FastSMC = True
    p.hashing = True
    for k, v in dict(REGRESSION_PARAMS, **kw).items():
        setattr(p, k, v)
    p.validateParamsFastSMC()
    return p


@pytest.mark.parametrize("job", [dict(), dict(jobs=9, jobInd=7, hashing=False)])
def test_data_and_model_tables_match_oracle(asmc, oracle_mod, job):
    """Data loading (job windows, folding, map interpolation), the RNG-driven emission tables and the per-site
    transition rows are bit-identical to the oracle's."""
    p = _params(asmc, **job)
    d = asmc.Data(p)
    # std::rand() is process-global state seeded by Data (srand(1234)): draw the emission tables before the oracle
    # instance reseeds and consumes it
    t = asmc.pyASMC.prepareModelTables(d, p)
    okw = dict(REGRESSION_PARAMS, hashing=job.get("hashing", True))
    o = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, "/tmp/x", jobs=job.get("jobs", 1),
                          jobInd=job.get("jobInd", 1), **okw)
    gen, phys = o.positions()
    assert d.sites == o.sites and 2 * len(d.IIDList) == o.num_haps
    assert np.array_equal(np.array(d.geneticPositions, np.float32).view(np.uint32), gen.view(np.uint32))
    assert np.array_equal(np.array(d.physicalPositions), phys)
    assert np.array_equal(np.array(d.siteWasFlippedDuringFolding, bool), o.flipped().astype(bool))
    from fastsmc_b200 import _native
    assert np.array_equal(d.hapBits, _native.pack_haplotypes(o.haplotypes()))
    assert (d.windowSize, d.w_i, d.w_j, d.is_j_above_diag) == (o.window_size, o.w_i, o.w_j, o.above_diag)

    e1, e0m1, e2m0 = o.emissions()
    for mine, ref in ((t["emission1"], e1), (t["emission0minus1"], e0m1), (t["emission2minus0"], e2m0)):
        assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))
    assert t["state_threshold"] == o.state_threshold and t["age_threshold"] == o.age_threshold
    assert np.float32(t["probability_threshold"]) == np.float32(o.probability_threshold)
    # transition rows: the row chosen for every gap equals the oracle's lookup by rounded distance
    rows = t["distance_row"]
    for s in (1, 2, 100, 3333, d.sites - 1):
        key = oracle_mod.round_morgans(float(gen[s] - gen[s - 1]))
        ref = o.transition(key)
        for name, r in zip(("D", "B", "U", "RR"), ref):
            assert np.array_equal(t[name][rows[s]].view(np.uint32), r.view(np.uint32))


def _brute_force_intervals(hap_bits, n_words, gap):
    """All match intervals (a<b, startWord, endWord) by the order-free definition, numpy only."""
    H = hap_bits.shape[0]
    out = []
    eq = hap_bits[:, None, :n_words] == hap_bits[None, :, :n_words]  # [H][H][W]
    ia, ib = np.triu_indices(H, 1)
    m = eq[ia, ib]  # [pairs][W]
    has = np.flatnonzero(m.any(axis=1))
    for p in has:
        ws = np.flatnonzero(m[p])
        s = e = ws[0]
        for w in ws[1:]:
            if w - e <= gap + 1:
                e = w
            else:
                out.append((ia[p], ib[p], s, e))
                s = e = w
        out.append((ia[p], ib[p], s, e))
    return np.array(out, np.int64)


def test_reference_candidate_order_replay(asmc, oracle_mod):
    """Order-free intervals + the host's replay of the reference's hash-map iteration order reproduce the oracle's
    decodeFromHashing call stream exactly (which itself reproduces golden G1's record order)."""
    p = _params(asmc)
    d = asmc.Data(p)
    o = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, "/tmp/x", hashing=True, **REGRESSION_PARAMS)
    want = o.seed()
    iv = _brute_force_intervals(d.hapBits, d.sites // 64, p.gap)
    for fast in (False, True):  # the literal replay of the two hash maps, and the production (sort-key) form
        order = asmc.pyASMC.replayReferenceOrder(iv, d, p.gap, p.min_m, fast=fast)
        got = iv[np.array(order)]
        got = np.stack([got[:, 0], got[:, 1], got[:, 2] * 64, got[:, 3] * 64 + 63], axis=1)
        assert len(got) == len(want) == 495
        assert np.array_equal(got, want.astype(np.int64))
    # independent of the order in which the intervals are handed over (the GPU emits them in no particular order)
    perm = np.random.default_rng(3).permutation(len(iv))
    order = asmc.pyASMC.replayReferenceOrder(iv[perm], d, p.gap, p.min_m, fast=True)
    assert np.array_equal(iv[perm][np.array(order)], iv[np.array(asmc.pyASMC.replayReferenceOrder(iv, d, p.gap, p.min_m))])


@pytest.mark.parametrize("n_haps,n_sites,gap,min_m", [(240, 3200, 1, 0.0), (400, 1920, 0, 0.05), (150, 6400, 2, 0.4)])
def test_fast_order_replay_equals_literal_replay_on_dense_matches(asmc, tmp_path, n_haps, n_sites, gap, min_m):
    """Synthetic data with few founders: hundreds of thousands of short intervals, most of which never become
    candidates but all of which are nodes of the reference's extend map; the map grows through many rehashes and
    buckets empty and refill.  The sort-key replay must give the literal replay's sequence."""
    from fastsmc_b200 import synth
    root = str(tmp_path / "dense")
    synth.dataset(root, n_haps, n_sites, 3000 * n_sites, 1, 77 + n_haps, founders=6)
    p = asmc.DecodingParams()
    p.verbose = False
    p.inFileRoot, p.decodingQuantFile, p.outFileRoot = root, DQ_69, root + ".out"
    p.decodingModeString, p.foldData, p.usingCSFS, p.FastSMC, p.hashing = "array", True, True, True, True
    p.gap, p.min_m = gap, min_m
    p.validateParamsFastSMC()
    d = asmc.Data(p)
    iv = _brute_force_intervals(np.array(d.hapBits), d.sites // 64, gap)  # folding flips both haplotypes alike
    assert len(iv) > 20000
    slow = asmc.pyASMC.replayReferenceOrder(iv, d, gap, min_m, fast=False)
    fast = asmc.pyASMC.replayReferenceOrder(iv, d, gap, min_m, fast=True)
    assert len(slow) > 100 and list(slow) == list(fast)
    # words that start very many intervals are sorted on all threads: same result when every word takes that path
    os.environ["FSMC_BIG_WORD"] = "1"
    try:
        assert list(asmc.pyASMC.replayReferenceOrder(iv, d, gap, min_m, fast=True)) == list(slow)
    finally:
        del os.environ["FSMC_BIG_WORD"]


def test_hmm_utils_known_answers(asmc):
    """ASMC_SRC/TESTS/test_hmm_utils.cpp:180-229, 298-353 and ASMC_SRC/TESTS/test_hashing.cpp:120-151."""
    m = asmc.pyASMC
    f = np.float32
    assert m.roundMorgans(0.4, 2, 0.5) == 0.5 and m.roundMorgans(-1.0, 7, 0.5) == 0.5
    assert [f(m.roundMorgans(0.123456, q, 1e-10)) for q in range(5)] == [f(0.1), f(0.12), f(0.123), f(0.1235), f(0.12346)]
    assert [m.roundPhysical(v, q) for v in (-1, 0, 1) for q in (0, 1, 2)] == [1] * 9
    assert [m.roundPhysical(123456, q) for q in range(6)] == [100000, 120000, 123000, 123500, 123460, 123456]
    gen = [0.12, 0.23, 0.34, 0.45, 0.56, 0.67]
    assert [m.getFromPosition(gen, 4, c) for c in (1, 21, 23, 30, 45, 60)] == [3, 2, 1, 1, 0, 0]
    assert [m.getToPosition(gen, 1, c) for c in (1, 10, 12, 30, 40, 60)] == [3, 3, 4, 5, 6, 6]
    assert [m.getToPosition(gen, 6, c) for c in (1e-6, 1, 10)] == [6, 6, 6]
    g9 = [0.1 * i for i in range(9)]
    assert m.cmBetween(1, 5, g9, 4) == pytest.approx(100.0 * (f(g9[8]) - f(g9[4])), rel=1e-6)
    assert m.hapToDipId(7) == (3, 2) and m.dipToHapId(3, 2) == 7
    assert m.indPlusHapToCombinedId("abc", 2) == "abc#2"
    assert m.combinedIdToIndPlusHap("1_10#1") == ("1_10", 1)
    with pytest.raises(RuntimeError):
        m.combinedIdToIndPlusHap("nohash")
    assert m.getIndIdxFromIdString(["a", "b", "c"], "c") == 2
    with pytest.raises(RuntimeError):
        m.getIndIdxFromIdString(["a"], "z")


def test_decoding_quantities_validation(asmc, tmp_path):
    """ASMC_SRC/TESTS/test_decoding_quantities.cpp:26-48."""
    with pytest.raises(RuntimeError, match="does not exist"):
        asmc.DecodingQuantities(str(tmp_path / "random_nonexistent_file.txt"))
    bad = tmp_path / "bad.txt"
    bad.write_text('this file does not start with "TransitionType"\n')
    with pytest.raises(RuntimeError, match="does not seem to contain the correct information"):
        asmc.DecodingQuantities(str(bad))
    dq = asmc.DecodingQuantities(FASTSMC_EXAMPLE_DQ)
    assert dq.states == 159 and dq.CSFSSamples == 100 and len(dq.Dvectors) == len(dq.rowRatioVectors) > 1000


def test_binary_data_reader_fixture(asmc):
    """ASMC_SRC/TESTS/test_binary_data_reader.cpp:47-88 — the reference's binary fixture (1 520 records)."""
    r = asmc.BinaryDataReader(os.path.join(GOLDEN, "binary_output.bibd.gz"))
    lines = []
    while r.moreLinesInFile():
        lines.append(r.getNextLine())
    assert len(lines) == 1520
    a, b = lines[0], lines[1]
    assert (a.ind1FamId, a.ind1Id, a.ind1Hap, a.ind2FamId, a.ind2Id, a.ind2Hap) == ("1_94", "1_94", 1, "1_104", "1_104", 1)
    assert (a.chromosome, a.ibdStart, a.ibdEnd) == (1, 8740, 1660011)
    assert a.lengthInCentimorgans == pytest.approx(1.86962, rel=1e-5) and a.ibdScore == pytest.approx(0.403475, rel=1e-5)
    assert a.postEst == pytest.approx(146.203, rel=1e-5) and a.mapEst == pytest.approx(24.9999, rel=1e-5)
    assert (b.chromosome, b.ibdStart, b.ibdEnd) == (1, 1679626, 1679626)
    assert b.lengthInCentimorgans == pytest.approx(0.0, abs=1e-5) and b.ibdScore == pytest.approx(0.0175673, rel=1e-5)
    assert b.postEst == pytest.approx(18029.8, rel=1e-5)
    d = asmc.IbdPairDataLine()
    d.lengthInCentimorgans, d.postEst, d.mapEst = 1.2, 2.3, 3.4
    assert d.toString() == "0_00\t0_00\t-1\t0_00\t0_00\t-1\t-1\t-1\t-1\t1.2\t-1\t2.3\t3.4"


def test_decoding_params_defaults_and_validation(asmc):
    """ASMC_SRC/TESTS/test_decoding_params.cpp / test_unit_decoding_params.py: FastSMC constructor defaults."""
    p = asmc.DecodingParams()
    assert (p.jobs, p.jobInd, p.batchSize, p.time, p.gap, p.min_m, p.hashing) == (1, 1, 64, 100, 1, 1.0, False)
    q = asmc.DecodingParams(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ)
    assert q.decodingMode == asmc.DecodingMode.arrayFolded and q.foldData and q.usingCSFS and not q.FastSMC


def test_job_data_cut_from_one_read_equals_reading_the_job(asmc):
    """runAllJobs reads the files once and cuts every job's two sample windows out of the result (Data.forJob): the
    same object as Data(params) reading the job from the files (ref: Data.cpp:62-80, 212-262)."""
    whole = asmc.Data(_params(asmc))
    for jobs in (4, 9):
        for job_ind in range(1, jobs + 1):
            p = _params(asmc, jobs=jobs, jobInd=job_ind)
            a, b = asmc.Data(p), asmc.Data.forJob(whole, p)
            assert a.IIDList == b.IIDList and a.FamIDList == b.FamIDList and len(a.IIDList) > 0
            assert list(a.globalHapId) == list(b.globalHapId)
            assert (a.windowSize, a.w_i, a.w_j, a.is_j_above_diag) == (b.windowSize, b.w_i, b.w_j, b.is_j_above_diag)
            assert np.array_equal(np.array(a.hapBits), np.array(b.hapBits))
            assert a.sites == b.sites and list(a.physicalPositions) == list(b.physicalPositions)
            assert np.array_equal(np.array(a.geneticPositions), np.array(b.geneticPositions))
            assert a.calculateUndistinguishedCounts(50) == b.calculateUndistinguishedCounts(50)


def test_packed_matrix_cache_equals_text_read(asmc, tmp_path):
    """Input codec: with hapBitCache the packed matrix of the whole data set is written next to the haps file
    on the first read and loaded instead of the gz text afterwards — for the whole data set and for every job cut out of
    it — and a cache whose source files changed is ignored."""
    import shutil
    import time
    from fastsmc_b200 import synth
    root = str(tmp_path / "syn")
    synth.dataset(root, 120, 1500, 4_500_000, 1, 5)

    def params(cache, jobs=1, job_ind=1):
        p = asmc.DecodingParams()
        p.verbose = False
        p.inFileRoot, p.decodingQuantFile, p.outFileRoot = root, DQ_69, root + ".out"
        p.decodingModeString, p.foldData, p.usingCSFS, p.FastSMC, p.hashing = "array", True, True, True, True
        p.jobs, p.jobInd, p.hapBitCache = jobs, job_ind, cache
        p.validateParamsFastSMC()
        return p

    def same(x, y):
        assert x.sites == y.sites and list(x.IIDList) == list(y.IIDList) and x.chrNumber == y.chrNumber
        for f in ("hapBits", "flipMask", "globalHapId", "totalSamplesCount", "derivedAlleleCounts", "geneticPositions",
                  "physicalPositions"):
            assert np.array_equal(np.array(getattr(x, f)), np.array(getattr(y, f))), f
        assert (x.windowSize, x.w_i, x.w_j, x.is_j_above_diag) == (y.windowSize, y.w_i, y.w_j, y.is_j_above_diag)

    cache = root + ".hap.gz.fsmcbits"
    plain = asmc.Data(params(False))
    assert not os.path.exists(cache)
    first = asmc.Data(params(True))       # reads the text, writes the cache
    assert os.path.exists(cache)
    second = asmc.Data(params(True))      # reads the cache
    same(plain, first)
    same(plain, second)
    for jobs, job_ind in ((4, 1), (4, 3), (4, 4), (9, 5)):
        same(asmc.Data(params(False, jobs, job_ind)), asmc.Data(params(True, jobs, job_ind)))
    # a changed haps file invalidates the cache: flip one allele of the first site and rewrite
    import gzip
    lines = gzip.open(root + ".hap.gz", "rt").read().splitlines()
    t = lines[0].split(" ")
    t[5] = "1" if t[5] == "0" else "0"
    lines[0] = " ".join(t)
    time.sleep(1.1)  # modification times have one-second resolution
    with gzip.open(root + ".hap.gz", "wt") as f:
        f.write("\n".join(lines) + "\n")
    changed_plain = asmc.Data(params(False))
    changed_cached = asmc.Data(params(True))
    same(changed_plain, changed_cached)
    assert not np.array_equal(np.array(changed_plain.hapBits), np.array(plain.hapBits))


def test_nested_seed_ranks_follow_the_sub_hash_partition(asmc, tmp_path):
    """max_seeds (SeedHash.hpp:56-69, 85-93): the seed-map rank of a haplotype at word w is the position of its FINAL nested
    bucket in the depth-first visit of the sub-hashes.  Checked here: two haplotypes share a rank iff a brute-force walk of
    the nesting rule (bucket larger than max_seeds and next word inside the read-ahead buffer -> split on the next word)
    puts them in the same final bucket; ranks are dense; without max_seeds the ranks are those of the words alone."""
    from fastsmc_b200 import synth
    root = str(tmp_path / "dense")
    synth.dataset(root, 60, 1280, 3000 * 1280, 1, 3, founders=4)
    p = _params(asmc)
    p.inFileRoot = root
    d = asmc.Data(p)
    H, W = d.hapBits.shape[0], d.sites // 64
    raw = (d.hapBits[:, :W] ^ np.asarray(d.flipMask, dtype=np.uint64)[None, :W])
    plain = asmc.pyASMC.seedGroupRanks(d)
    for max_seeds, read_ahead in ((5, 10), (12, 3), (1, 10)):
        ranks = asmc.pyASMC.seedGroupRanks(d, max_seeds=max_seeds, read_ahead=read_ahead)
        assert ranks.shape == (W, H)
        deeper = 0
        for w in range(W):
            read_words = min(W, w + read_ahead)
            final = {}  # haplotype -> key of its final bucket
            stack = [(w, np.arange(H))]
            while stack:
                level, members = stack.pop()
                for value in np.unique(raw[members, level]):
                    bucket = members[raw[members, level] == value]
                    if len(bucket) > max_seeds and level + 1 < read_words:
                        stack.append((level + 1, bucket))
                        deeper += 1
                    else:
                        for h in bucket:
                            final[h] = (level, int(bucket[0]))
            keys = {}
            for h in range(H):
                keys.setdefault(final[h], set()).add(int(ranks[w, h]))
            assert all(len(v) == 1 for v in keys.values())                      # one rank per final bucket
            assert len({next(iter(v)) for v in keys.values()}) == len(keys)      # different buckets, different ranks
            assert sorted(set(ranks[w].tolist())) == list(range(len(keys)))     # dense
        assert deeper > 0
    assert np.array_equal(asmc.pyASMC.seedGroupRanks(d, max_seeds=H + 1), plain)
