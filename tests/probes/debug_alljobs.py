"""Developer probe: runAllJobs (J=4 jobs of the example data) on one or two host threads, exact arithmetic, every job's
file diffed against the oracle's.  Usage: python tests/probes/debug_alljobs.py [repeats]"""
import gzip
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, REGRESSION_PARAMS  # noqa: E402
from fastsmc_b200 import asmc  # noqa: E402
from oracle import pyoracle  # noqa: E402

pyoracle.build()
repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 2


def lines(path):
    with gzip.open(path, "rt") as f:
        return f.read().splitlines()


tmp = tempfile.mkdtemp()
want = {}
for j in range(1, 5):
    o = pyoracle.Oracle(FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ, os.path.join(tmp, "o"), hashing=True, jobs=4, jobInd=j,
                        **REGRESSION_PARAMS)
    path = os.path.join(tmp, f"oracle{j}.ibd.gz")
    o.run(path)
    want[j] = lines(path)

for devices in ([0], [0, 0], [0, 0, 0, 0]):
    for rep in range(repeats):
        p = asmc.DecodingParams()
        p.verbose = False
        p.inFileRoot, p.decodingQuantFile = FASTSMC_EXAMPLE, FASTSMC_EXAMPLE_DQ
        p.outFileRoot = os.path.join(tmp, f"gpu_{len(devices)}_{rep}")
        p.decodingModeString, p.foldData, p.usingCSFS, p.FastSMC, p.hashing = "array", True, True, True, True
        for k, v in dict(REGRESSION_PARAMS, exactArithmetic=True, jobs=4).items():
            setattr(p, k, v)
        p.validateParamsFastSMC()
        reports = asmc.pyASMC.runAllJobs(p, devices)
        for r in reports:
            got = lines(f"{p.outFileRoot}.{r.jobInd}.4.FastSMC.ibd.gz")
            w = want[r.jobInd]
            status = "OK" if got == w else "DIFF"
            print(f"threads={len(devices)} rep={rep} job={r.jobInd} dev={r.device} err={r.error!r} got={len(got)} "
                  f"want={len(w)} cand={r.candidates} {status}", flush=True)
            if got != w:
                sg, sw = set(got), set(w)
                print("  only in got:", len(sg - sw), " only in want:", len(sw - sg), " same multiset:",
                      sorted(got) == sorted(w))
                for x in sorted(sg - sw)[:4]:
                    print("   +", x)
                for x in sorted(sw - sg)[:4]:
                    print("   -", x)
                for i, (a, b) in enumerate(zip(got, w)):
                    if a != b:
                        print("  first diff at", i)
                        print("   got ", got[max(0, i - 1):i + 3])
                        print("   want", w[max(0, i - 1):i + 3])
                        break
