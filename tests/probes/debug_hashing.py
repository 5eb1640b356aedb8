"""Developer probe: FastSMC.run() with hashing on the example data, exact arithmetic, repeated; the candidate list and
the output lines of every repetition are compared with the oracle's.  Usage: python tests/probes/debug_hashing.py [reps]"""
import gzip
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, REGRESSION_PARAMS  # noqa: E402
from fastsmc_b200 import asmc  # noqa: E402
from oracle import pyoracle  # noqa: E402

pyoracle.build()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
exact = os.environ.get("PROBE_EXACT", "1") == "1"
tmp = tempfile.mkdtemp()


def lines(path):
    with gzip.open(path, "rt") as f:
        return f.read().splitlines()


o = pyoracle.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, os.path.join(tmp, "o"), hashing=True, **REGRESSION_PARAMS)
ref = os.path.join(tmp, "oracle.ibd.gz")
o.run(ref)
want = lines(ref)
want_c = o.candidates()
print("oracle:", len(want), "lines,", len(want_c), "candidates", flush=True)
bad = 0
for rep in range(reps):
    p = asmc.DecodingParams()
    p.verbose = False
    p.inFileRoot, p.decodingQuantFile = FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ
    p.outFileRoot = os.path.join(tmp, f"gpu_{rep}")
    p.decodingModeString, p.foldData, p.usingCSFS, p.FastSMC, p.hashing = "array", True, True, True, True
    for k, v in dict(REGRESSION_PARAMS, exactArithmetic=exact).items():
        setattr(p, k, v)
    p.validateParamsFastSMC()
    f = asmc.FastSMC(p)
    f.setKeepCandidates(True)
    f.run()
    got = lines(f"{p.outFileRoot}.{p.jobInd}.{p.jobs}.FastSMC.ibd.gz")
    c = np.asarray(f.getCandidates()).reshape(-1, 4)
    same_c = c.shape == want_c.shape and bool((c == want_c).all())
    ndiff = sum(a != b for a, b in zip(got, want)) + abs(len(got) - len(want))
    print(f"rep {rep}: candidates {'same' if same_c else 'DIFFER'} lines differing {ndiff}", flush=True)
    if not same_c:
        rows = np.nonzero((c != want_c).any(axis=1))[0] if c.shape == want_c.shape else []
        for r in rows[:5]:
            print("   cand", r, "got", c[r], "want", want_c[r])
    if ndiff:
        bad += 1
        for i, (a, b) in enumerate(zip(got, want)):
            if a != b:
                print("   line", i, "\n     got ", a, "\n     want", b)
                break
print("bad repetitions:", bad, "of", reps)
