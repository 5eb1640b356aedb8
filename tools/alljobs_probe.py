"""Developer probe (not the bench): all J jobs of one synthetic data set dealt to G GPUs of this box by
pyASMC.runAllJobs (one host thread and one context per GPU, jobs from a shared longest-first queue) — the strong-scaling
form of SURVEY §8(e).

    python tools/alljobs_probe.py [n_diploid] [n_sites] [span_mb] [jobs] [gpus] [reference_order 0|1] [out.json]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastsmc_b200 import asmc, synth  # noqa: E402

n_dip = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
n_sites = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
span_mb = float(sys.argv[3]) if len(sys.argv) > 3 else 240.0
jobs = int(sys.argv[4]) if len(sys.argv) > 4 else 16
gpus = int(sys.argv[5]) if len(sys.argv) > 5 else 1
ref_order = bool(int(sys.argv[6])) if len(sys.argv) > 6 else False
out_json = sys.argv[7] if len(sys.argv) > 7 else None

root = f"/tmp/fsmc_scale/d{n_dip}_s{n_sites}_c1"
if not os.path.exists(root + ".hap.gz"):
    synth.dataset(root, 2 * n_dip, n_sites, int(span_mb * 1e6), 1, 20201117 + 4)
p = asmc.DecodingParams()
p.verbose = False
p.inFileRoot, p.decodingQuantFile, p.outFileRoot = root, os.path.join(ROOT, "data", "30-100-2000.decodingQuantities.gz"), root + ".alljobs"
p.decodingModeString, p.foldData, p.usingCSFS = "array", True, True
p.FastSMC, p.hashing, p.batchSize, p.time = True, True, 32, 50
p.min_m, p.gap = 1.5, 1
p.doPerPairMAP = p.doPerPairPosteriorMean = p.outputIbdSegmentLength = True
p.useKnownSeed = True
p.referenceCandidateOrder = ref_order
p.jobs = jobs
p.validateParamsFastSMC()
t0 = time.perf_counter()
reports = asmc.pyASMC.runAllJobs(p, list(range(gpus)))
wall = time.perf_counter() - t0
rep = {"n_diploid": n_dip, "sites": n_sites, "jobs": jobs, "gpus": gpus, "reference_order": ref_order, "wall_s": wall,
       "errors": [r.error for r in reports if r.error], "segments": sum(r.segments for r in reports),
       "candidates": sum(r.candidates for r in reports), "pair_sites": sum(r.pairSites for r in reports),
       "decode_kernel_ms": sum(r.kernelMs for r in reports), "seed_kernel_ms": sum(r.seedMs for r in reports),
       "job_wall_s": [round(r.wallSeconds, 2) for r in reports], "job_device": [r.device for r in reports],
       "job_stages_s": [{k: round(getattr(r, k), 3) for k in ("prepareSeconds", "tablesSeconds", "uploadSeconds", "seedSeconds",
                                                             "orderSeconds", "decodeSeconds", "outputSeconds") if hasattr(r, k)}
                        for r in reports],
       "busy_s_per_gpu": {d: round(sum(r.wallSeconds for r in reports if r.device == d), 2) for d in range(gpus)}}
print(json.dumps(rep, indent=1))
if out_json:
    os.makedirs(os.path.dirname(os.path.abspath(out_json)), exist_ok=True)
    json.dump(rep, open(out_json, "w"), indent=1)
