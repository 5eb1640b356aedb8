#!/usr/bin/env python
"""How long one 'round' of tiles takes as a function of the resident warps per SM (cfg2 tiles, default flags, narrow kernel):
the number behind the strong-scaling discussion of DESIGN.md §6.  python tools/round_probe.py [tile counts ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fastsmc_b200 import _native as N, asmc  # noqa: E402


def main():
    counts = [int(x) for x in sys.argv[1:]] or [148, 296, 592, 888, 1184, 1776, 1951, 2368]
    root = bench.make_dataset(0)
    p = asmc.DecodingParams()
    p.verbose = False
    p.inFileRoot, p.decodingQuantFile, p.outFileRoot = root, bench.DQ, "/tmp/fsmc_bench/r0/out"
    p.decodingModeString, p.foldData, p.usingCSFS = "array", True, True
    p.FastSMC, p.hashing, p.batchSize, p.time = True, False, 32, 50
    p.doPerPairMAP = p.doPerPairPosteriorMean = p.outputIbdSegmentLength = True
    p.useKnownSeed = True
    p.validateParamsFastSMC()
    data = asmc.Data(p)
    tables = asmc.pyASMC.prepareModelTables(data, p)
    L = data.sites
    a, b = bench.all_pairs_in_reference_order(len(data.IIDList))
    tiles = bench.tiles_for(a, b, L)
    ctx = N.Context(0)
    st = torch.cuda.Stream()
    ctx.set_stream(st.cuda_stream)
    ctx.set_model(**dict(tables, age_threshold=tables["state_threshold"]))
    ctx.set_haplotypes(data.hapBits, L)
    for T in counts:
        sub = {k: (v[:T] if k != "rows" else v) for k, v in tiles.items()}
        pl = ctx.plan(sub, N.CALL_SEGMENTS | N.SEG_AGE, segment_capacity=1 << 21)
        pl.launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3):
            pl.launch()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        pl.collect()
        pl.close()
        print(f"tiles {T:5d}  {T / 148:5.2f} per SM  {ms:8.2f} ms  {T * 32 * L / ms / 1e6:8.2f} G pair-sites/s", flush=True)


if __name__ == "__main__":
    main()
