"""Developer probe (CPU only): the reference-order replay on synthetic intervals at scale, with its phase timings
(FSMC_TRACE=1).  Usage: FSMC_TRACE=1 python tests/probes/replay_scale.py [n_intervals] [n_haps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fastsmc_b200 import asmc, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
W = 781
root = f"/tmp/fsmc_scale/replay_{H}"
if not os.path.exists(root + ".hap.gz"):
    synth.dataset(root, H, 50000, 240_000_000, 1, 1)
p = asmc.DecodingParams()
p.verbose = False
p.inFileRoot, p.decodingQuantFile, p.outFileRoot = root, os.path.join(ROOT, "data", "30-100-2000.decodingQuantities.gz"), root + ".out"
p.decodingModeString, p.foldData, p.usingCSFS = "array", True, True
p.FastSMC, p.hashing, p.batchSize, p.time = True, True, 32, 50
p.validateParamsFastSMC()
d = asmc.Data(p)
rng = np.random.default_rng(1)
a = rng.integers(0, H - 1, n)
b = rng.integers(1, H, n)
lo, hi = np.minimum(a, b), np.maximum(a, b)
hi = np.where(hi == lo, hi + 1, hi)
start = rng.integers(0, W, n)
length = np.minimum(rng.geometric(0.5, n) - 1 + (rng.random(n) < 0.1) * rng.integers(3, 12, n), W - 1 - start)
iv = np.stack([lo, hi, start, start + length], axis=1).astype(np.int64)
t = time.time()
order = asmc.pyASMC.replayReferenceOrder(iv, d, 1, 1.5, fast=True)
print(n, "intervals, fast replay incl. binding conversion %.2f s" % (time.time() - t), "emitted", len(order))
