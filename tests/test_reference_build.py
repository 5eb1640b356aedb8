"""The oracle pinned to the REFERENCE ITSELF: the reference's own sources (HMM.cpp, FastSMC.cpp, HASHING/*, Data.cpp, ...)
compiled unmodified against the shim headers of oracle/shim (oracle/Makefile) are run here on the bundled example, and the
oracle restatement must write the same .ibd.gz, line for line — in the NO_SSE flavour (exact 1/x) and in the reference's
default AVX flavour (approximate reciprocal), with hashing (seeding + boost node order + batching + caller) and without
(job slice of the all-pairs enumeration).  The reference's committed goldens were produced with an older libstdc++
(std::shuffle differs), which is why this build is compared with the oracle and not with the golden files;
the oracle reproduces those too (test_oracle_golden.py)."""
import gzip
import os

import pytest

from conftest import FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, REGRESSION_PARAMS


def _lines(path):
    with gzip.open(path, "rb") as f:
        return f.read().splitlines()


@pytest.mark.parametrize("flavour,simd", [("nosse", False), ("avx", True)])
@pytest.mark.parametrize("hashing,jobs,job_ind", [(True, 1, 1), (True, 4, 3), (False, 25, 8)], ids=["hashing", "hashing-job3of4", "allpairs-job8of25"])
def test_oracle_equals_reference_build(oracle_mod, tmp_path, flavour, simd, hashing, jobs, job_ind):
    if oracle_mod.reference_binary(flavour) is None:
        pytest.skip("reference build not present (needs /root/reference at build time)")
    out = str(tmp_path / "ref")
    oracle_mod.reference_run(flavour, FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, out, hashing=hashing, jobs=jobs, jobInd=job_ind)
    ref = _lines(f"{out}.{job_ind}.{jobs}.FastSMC.ibd.gz")
    o = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, str(tmp_path / "orc"), hashing=hashing, jobs=jobs,
                          jobInd=job_ind, shuffleFlavor=0, simdFlavor=simd, **REGRESSION_PARAMS)
    path = str(tmp_path / "oracle.ibd.gz")
    o.run(path)
    mine = _lines(path)
    assert len(ref) > 100
    assert mine == ref


def test_reference_default_flags_conditional_age_estimates(oracle_mod, tmp_path):
    """FastSMC_exe's default age estimates (conditional on TMRCA < time) through the reference build vs the oracle."""
    if oracle_mod.reference_binary("nosse") is None:
        pytest.skip("reference build not present")
    out = str(tmp_path / "ref")
    oracle_mod.reference_run("nosse", FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, out, hashing=True, noConditionalAgeEstimates=False)
    params = dict(REGRESSION_PARAMS, noConditionalAgeEstimates=False)
    o = oracle_mod.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, str(tmp_path / "orc"), hashing=True, shuffleFlavor=0, **params)
    path = str(tmp_path / "oracle.ibd.gz")
    o.run(path)
    assert _lines(path) == _lines(f"{out}.1.1.FastSMC.ibd.gz")


@pytest.mark.parametrize("options", [dict(skip=0.13), dict(skip=0.145, gap=2, min_m=2.5), dict(max_seeds=20), dict(max_seeds=8, skip=0.13)],
                         ids=["skip0.13", "skip0.145-gap2", "max_seeds20", "max_seeds8-skip0.13"])
def test_oracle_equals_reference_build_with_skip_and_max_seeds(oracle_mod, tmp_path, options):
    """The non-default seeding options (low-complexity word skipping, sub-hashing of oversized buckets; FastSMC.cpp:208-219,
    HASHING/SeedHash.hpp:56-69, 85-93) on dense synthetic data (6 founders: a quarter to a half of the words are low
    complexity at these thresholds, buckets of 20-60 haplotypes): reference build vs the oracle, whole .ibd.gz."""
    if oracle_mod.reference_binary("nosse") is None:
        pytest.skip("reference build not present")
    from fastsmc_b200 import synth
    root = str(tmp_path / "dense")
    synth.dataset(root, 200, 1920, 3000 * 1920, 1, 11, founders=6)
    out = str(tmp_path / "ref")
    options = dict(dict(min_m=2.0), **options)
    oracle_mod.reference_run("nosse", root, FASTSMC_EXAMPLE_DQ, out, hashing=True, **options)
    ref = _lines(f"{out}.1.1.FastSMC.ibd.gz")
    params = dict(REGRESSION_PARAMS, **options)
    o = oracle_mod.Oracle(root, FASTSMC_EXAMPLE_DQ, str(tmp_path / "orc"), hashing=True, shuffleFlavor=0, **params)
    path = str(tmp_path / "oracle.ibd.gz")
    o.run(path)
    assert len(ref) > 500
    assert _lines(path) == ref
