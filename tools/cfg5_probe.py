"""Dry run of BASELINE.json configs[4] on ONE B200: 487 409 synthetic diploid samples x 11 528 chr22 array SNPs.

    python tools/cfg5_probe.py [out.json] [n_diploid] [jobs] [n_pairs]

Stages (seconds, one GPU, this box's host cores):
  generate   synthetic haplotypes written as the packed-matrix cache (fastsmc_b200.synth.packed_dataset; no gz text)
  read       Data(params) from the cache (the whole data set: 974 818 haplotypes x 181 words = 1.4 GB)
  tables     emission / transition tables of the whole data set (the reference's RNG-exact undistinguished counts: 3 x 11 528
             shuffles of 974 816 shorts, drawn once and shared by every job cut out of the data set)
  jobs       a SAMPLE of the J = 256 jobs of the reference's jobs/jobInd partition (the last job, two off-diagonal, one
             diagonal), each ASMC::FastSMC(params, Data::forJob(whole)).run(): seeding + reference candidate order on the
             device, decoding, .ibd.gz; the per-stage seconds of the sample are scaled to all 256 jobs on 8 GPUs
  per_site   per-site posterior mean TMRCA of a fixed sample of 10^6 haplotype pairs through the C ABI (fsmc_decode with
             FSMC_SITE_MEAN, host buffers) — the device path of ASMC::decodePairs(per_pair_posterior_means=True)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastsmc_b200 import _native as N, asmc, synth  # noqa: E402

out_json = sys.argv[1] if len(sys.argv) > 1 else None
n_dip = int(sys.argv[2]) if len(sys.argv) > 2 else 487_409
J = int(sys.argv[3]) if len(sys.argv) > 3 else 256
n_pairs = int(sys.argv[4]) if len(sys.argv) > 4 else 1_000_000
SITES, SPAN, CHROM, SEED = 11_528, 35_000_000, 22, 20201117 + 5
DQ = os.path.join(ROOT, "data", "30-100-2000.decodingQuantities.gz")
root = f"/tmp/fsmc_scale/cfg5_{n_dip}"
rep = {"n_diploid": n_dip, "sites": SITES, "jobs": J, "host_cores": os.cpu_count()}

t0 = time.perf_counter()
if not os.path.exists(root + ".hap.gz.fsmcbits"):
    synth.packed_dataset(root, 2 * n_dip, SITES, SPAN, CHROM, SEED)
rep["generate_s"] = time.perf_counter() - t0
rep["cache_bytes"] = os.path.getsize(root + ".hap.gz.fsmcbits")


def params(jobs=1, job_ind=1):
    p = asmc.DecodingParams()
    p.verbose = False
    p.inFileRoot, p.decodingQuantFile, p.outFileRoot = root, DQ, root + ".out"
    p.decodingModeString, p.foldData, p.usingCSFS = "array", True, True
    p.FastSMC, p.hashing, p.batchSize, p.time, p.min_m, p.gap = True, True, 32, 50, 1.5, 1
    p.doPerPairMAP = p.doPerPairPosteriorMean = p.outputIbdSegmentLength = True
    p.useKnownSeed, p.hapBitCache = True, True
    p.jobs, p.jobInd = jobs, job_ind
    p.validateParamsFastSMC()
    return p


t0 = time.perf_counter()
whole = asmc.Data(params())
rep["read_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
tables = asmc.pyASMC.prepareModelTables(whole, params())
rep["tables_s"] = time.perf_counter() - t0

# ---- a sample of the jobs -------------------------------------------------------------------------------------------------
side = int(round(J ** 0.5))
diagonal = [j for j in range(1, J + 1) if int(j ** 0.5) ** 2 == j and j != J]
sample = [J, 2, J - 2, diagonal[len(diagonal) // 2]]
it = iter(sample)
t0 = time.perf_counter()
reports = asmc.pyASMC.runJobs(params(J, 1), whole, lambda: next(it, 0), [0, 0])
rep["jobs_sample_wall_s"] = time.perf_counter() - t0
rep["jobs_sample"] = []
for r in reports:
    if r.error:
        raise RuntimeError(r.error)
    rep["jobs_sample"].append({k: getattr(r, k) for k in (
        "jobInd", "candidates", "pairsDecoded", "segments", "pairSites", "kernelMs", "seedMs", "wallSeconds", "prepareSeconds",
        "tablesSeconds", "uploadSeconds", "seedSeconds", "orderSeconds", "decodeSeconds", "outputSeconds")})
off = [r for r in rep["jobs_sample"] if r["jobInd"] in (2, J - 2)]
dia = [r for r in rep["jobs_sample"] if r["jobInd"] in diagonal]
n_off, n_dia = J - side, side  # J = side^2 jobs: side diagonal ones (the last job among them), the rest off-diagonal
stages = ("prepareSeconds", "seedSeconds", "decodeSeconds", "outputSeconds")
mean = lambda rows, k: sum(r[k] for r in rows) / max(len(rows), 1)
all_jobs = {k: n_off * mean(off, k) + n_dia * mean(dia, k) for k in stages + ("wallSeconds", "pairSites", "candidates", "segments")}
all_jobs["kernel_s"] = (n_off * mean(off, "kernelMs") + n_dia * mean(dia, "kernelMs") + n_off * mean(off, "seedMs") + n_dia * mean(dia, "seedMs")) / 1e3
rep["all_jobs_extrapolated"] = dict(all_jobs, note="sum over the 256 jobs of the sampled jobs' stage seconds (off-diagonal and diagonal jobs "
                                                   "scaled separately); two jobs run concurrently per GPU")
rep["eight_gpus_estimate_s"] = {"jobs_wall": all_jobs["wallSeconds"] / 2 / 8, "kernels": all_jobs["kernel_s"] / 8}

# ---- per-site posterior means of 10^6 pairs ----------------------------------------------------------------------------------
ctx = N.Context(0)
t0 = time.perf_counter()
ctx.set_model(**dict(tables, age_threshold=tables["emission1"].shape[1]))
ctx.set_haplotypes(whole.hapBits, SITES)
rep["per_site_upload_s"] = time.perf_counter() - t0
rng = np.random.default_rng(SEED)
H = 2 * n_dip
a = rng.integers(0, H, n_pairs).astype(np.uint32)
b = ((a + 1 + rng.integers(0, H - 1, n_pairs)) % H).astype(np.uint32)
chunk = 32768
kernel_ms, pair_sites, checksum = 0.0, 0.0, 0.0
t0 = time.perf_counter()
for lo in range(0, n_pairs, chunk):
    hi = min(n_pairs, lo + chunk)
    tiles = ctx.make_tiles(a[lo:hi], b[lo:hi], sites=SITES)
    r = ctx.decode(tiles, N.SITE_MEAN)
    kernel_ms += r.stats.kernelMs
    pair_sites += r.stats.pairSites
    checksum += float(r.site_mean[tiles["rows"]].mean())
rep["per_site"] = {"pairs": n_pairs, "wall_s": time.perf_counter() - t0, "kernel_s": kernel_ms / 1e3, "pair_sites": pair_sites,
                   "pair_sites_per_s_kernel": pair_sites / (kernel_ms / 1e3), "output_bytes": 4.0 * n_pairs * SITES,
                   "mean_of_chunk_means_generations": checksum / ((n_pairs + chunk - 1) // chunk)}
# whole run on 8 GPUs: every rank reads the cache and draws the tables once; the jobs' stage seconds are spread over 8 GPUs x 2
# host threads, the per-site sample over 8 GPUs
parts = {"read": rep["read_s"], "tables": rep["tables_s"], "per_site": rep["per_site"]["wall_s"] / 8}
for k in stages:
    parts["jobs_" + k.replace("Seconds", "")] = all_jobs[k] / 16
total = sum(parts.values())
rep["eight_gpus_estimate_s"]["whole_run"] = total
rep["eight_gpus_estimate_s"]["stage_seconds"] = parts
rep["eight_gpus_estimate_s"]["stage_share"] = {k: v / total for k, v in parts.items()}
print(json.dumps(rep, indent=1))
if out_json:
    os.makedirs(os.path.dirname(os.path.abspath(out_json)), exist_ok=True)
    json.dump(rep, open(out_json, "w"), indent=1)
