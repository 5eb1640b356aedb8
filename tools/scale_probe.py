"""Developer probe (not the bench): hashing + decoding on a synthetic data set of a BASELINE.json shape, with a phase
breakdown (generate, read, model tables, seeding, candidate order, decode, output).

    python tools/scale_probe.py [n_diploid] [n_sites] [span_mb] [chrom] [jobs] [reference_order 0|1] [out.json]

cfg3 = 10000 50000 240 1 ; cfg4 = 100000 18022 64 20 (jobs 16/64) ; cfg5 = 487409 11528 35 22
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastsmc_b200 import asmc, synth  # noqa: E402

n_dip = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_sites = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
span_mb = float(sys.argv[3]) if len(sys.argv) > 3 else 240.0
chrom = int(sys.argv[4]) if len(sys.argv) > 4 else 1
jobs = int(sys.argv[5]) if len(sys.argv) > 5 else 1
ref_order = bool(int(sys.argv[6])) if len(sys.argv) > 6 else False
out_json = sys.argv[7] if len(sys.argv) > 7 else None
DQ = os.path.join(ROOT, "data", "30-100-2000.decodingQuantities.gz")

root = f"/tmp/fsmc_scale/d{n_dip}_s{n_sites}_c{chrom}"
rep = {"n_diploid": n_dip, "sites": n_sites, "span_mb": span_mb, "chrom": chrom, "jobs": jobs, "reference_order": ref_order}
t0 = time.perf_counter()
if not os.path.exists(root + ".hap.gz"):
    synth.dataset(root, 2 * n_dip, n_sites, int(span_mb * 1e6), chrom, 20201117 + 3)
rep["generate_s"] = time.perf_counter() - t0
rep["hap_gz_bytes"] = os.path.getsize(root + ".hap.gz")

p = asmc.DecodingParams()
p.verbose = False
p.inFileRoot, p.decodingQuantFile, p.outFileRoot = root, DQ, root + ".out"
p.decodingModeString, p.foldData, p.usingCSFS = "array", True, True
p.FastSMC, p.hashing, p.batchSize, p.time = True, True, 32, 50
p.min_m, p.gap = 1.5, 1
p.doPerPairMAP = p.doPerPairPosteriorMean = p.outputIbdSegmentLength = True
p.noConditionalAgeEstimates = bool(int(os.environ.get("FSMC_PROBE_UNCONDITIONAL", "0")))
p.useKnownSeed = True
p.referenceCandidateOrder = ref_order
p.jobs, p.jobInd = jobs, int(os.environ.get("FSMC_PROBE_JOBIND", "1"))
p.validateParamsFastSMC()

t0 = time.perf_counter()
data = asmc.Data(p)
rep["read_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
f = asmc.FastSMC(p, data)
rep["construct_s"] = time.perf_counter() - t0  # decoding quantities + emission tables + upload
t0 = time.perf_counter()
f.run()
rep["run_s"] = time.perf_counter() - t0
ss, st = f.getSeedingStats(), f.hmm().getRunStats()
rep.update(seed_wall_s=ss.seedWallS, order_wall_s=ss.orderWallS, submit_wall_s=ss.submitWallS, candidates=ss.candidates,
           intervals=ss.device.numIntervals, order_device_ms=ss.device.orderMs, order_epochs=ss.device.orderEpochs,
           rank_host_ms=ss.device.rankHostMs, max_live_nodes=ss.device.maxLiveNodes,
           seed_kernel_ms=ss.device.kernelMs, seed_launches=ss.device.kernelLaunches, pair_visits=ss.device.pairVisits,
           seed_starts=ss.device.numStarts, seed_matches=ss.device.numMatches, seed_bytes=ss.device.bytesRead,
           pairs_decoded=st.pairsDecoded, batches=st.batches, segments=st.segments, pair_sites=st.pairSites,
           decode_kernel_ms=st.kernelMs, decode_device_ms=st.deviceMs, decode_wall_s=st.decodeWallS,
           output_wall_s=st.outputWallS, decode_calls=st.decodeCalls)
if st.kernelMs > 0:
    rep["decode_pair_sites_per_s"] = st.pairSites / (st.kernelMs / 1e3)
if ss.device.kernelMs > 0:
    rep["seed_gb_per_s"] = ss.device.bytesRead / (ss.device.kernelMs / 1e3) / 1e9
print(json.dumps(rep, indent=1))
if out_json:
    os.makedirs(os.path.dirname(os.path.abspath(out_json)), exist_ok=True)
    json.dump(rep, open(out_json, "w"), indent=1)
