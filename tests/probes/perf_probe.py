"""Developer probe (not the bench): times the decode kernels on a synthetic all-pairs job, with model tables
taken from the oracle.  Usage: python tests/probes/perf_probe.py [n_haps] [n_sites] [dq: 69|159] [n_pairs]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import DQ_69, FASTSMC_EXAMPLE_DQ, context_from_oracle  # noqa: E402
from fastsmc_b200 import _native as N, synth  # noqa: E402
from oracle import pyoracle  # noqa: E402

n_haps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_sites = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
dq = DQ_69 if (len(sys.argv) <= 3 or sys.argv[3] == "69") else FASTSMC_EXAMPLE_DQ
n_pairs = int(sys.argv[4]) if len(sys.argv) > 4 else 0
flags_extra = int(sys.argv[5], 0) if len(sys.argv) > 5 else 0
root = f"/tmp/fsmc_probe/syn_{n_haps}_{n_sites}"
t = time.time()
if not os.path.exists(root + ".hap.gz"):
    synth.dataset(root, n_haps, n_sites, 3000 * n_sites, 1, 20201119)
print("synth", time.time() - t)
t = time.time()
cond = os.environ.get("PROBE_CONDITIONAL", "0") == "1"  # FastSMC's default: age estimates conditional on TMRCA < time
o = pyoracle.Oracle(root, dq, "/tmp/fsmc_probe/out", hashing=False, time=50, noConditionalAgeEstimates=not cond,
                    doPerPairMAP=True, doPerPairPosteriorMean=True)
print("oracle load", time.time() - t, o.sites, o.states, o.num_haps)
ctx = context_from_oracle(o, pyoracle)
ia, ib = np.triu_indices(n_haps, 1)
if n_pairs:
    ia, ib = ia[:n_pairs], ib[:n_pairs]
tiles = ctx.make_tiles(ia.astype(np.uint32), ib.astype(np.uint32), sites=o.sites) if len(ia) < 200000 else None
if tiles is None:
    # fast vectorised tiling for big jobs
    n = len(ia)
    T = (n + 31) // 32
    A = np.zeros(T * 32, np.uint32); B = np.zeros(T * 32, np.uint32)
    A[:n] = ia; B[:n] = ib
    tp = np.full(T, 32, np.int32); tp[-1] = n - 32 * (T - 1)
    z = np.zeros(T, np.int32); e = np.full(T, o.sites, np.int32)
    tiles = dict(hapA=A.reshape(T, 32), hapB=B.reshape(T, 32), tilePairs=tp, tileFrom=z, tileTo=e, tileScanFrom=z,
                 tileScanTo=e, rows=np.arange(n))
for flags, name in ((N.CALL_SEGMENTS | N.SEG_AGE, "segments+age"), (N.CALL_SEGMENTS, "segments"),
                    (N.CALL_SEGMENTS | N.SEG_AGE | N.EXACT, "segments+age exact")):
    if (flags & N.EXACT) and os.environ.get("PROBE_SKIP_EXACT", "0") == "1":
        continue
    flags |= flags_extra
    plan = ctx.plan(tiles, flags, segment_capacity=1 << 22)
    for it in range(3):
        plan.launch()
        r = plan.collect()
        ps = r.stats.pairSites
        print(f"{name}: kernel {r.stats.kernelMs:.2f} ms  {ps / r.stats.kernelMs / 1e6:.3f} G pair-sites/s  "
              f"segs {r.stats.numSegments} scratch {r.stats.scratchBytes / 2**30:.1f} GiB  S_kernel {r.stats.statesKernel} "
              f"beta GB/s {ps * o.states * 8 / r.stats.kernelMs / 1e6:.0f} narrow {r.stats.narrowKernel} tileWarps {r.stats.tileWarps}")
    plan.close()
