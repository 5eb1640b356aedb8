#!/usr/bin/env python
"""Benchmark of the FastSMC IBD hot path on B200 (BASELINE.json configs[1]: all-pairs decoding with hashing off,
1,000 synthetic haplotypes x 10,000 SNPs, 69-state decoding quantities; every pair is decoded over the whole sequence and
IBD segments with per-segment age estimates are called, as FastSMC::run does with hashing off).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over the whole job (499,500 haplotype pairs x 10,000 sites).
  value : pair-sites/s with the pair list resident in HBM (fsmc_plan_launch only inside the timed region)
  e2e   : the same through the C-ABI call fsmc_decode with HOST buffers: pair lists copied H2D and the segment records
          copied D2H inside the timed region
Under torchrun each rank owns one GPU and decodes one job of that size (jobs are independent: no collective on the data
path); value = all ranks' pair-sites / max-over-ranks device time  ("scaling": "weak").

--impl reference times the CPU restatement of the reference algorithm (oracle/, multi-threaded over batches, AVX2 lane
vectorisation like the reference's SIMD build) on a bounded sample of the same workload on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_HAPS, N_SITES, SPAN_BP, CHROM, SEED = 1000, 10_000, 30_000_000, 1, 20201117 + 2
DQ = os.path.join(ROOT, "data", "30-100-2000.decodingQuantities.gz")
FLOPS_PER_PAIR_SITE_STATE = 35  # 15 forward + 15 backward + 3 combine + 2 consume
NCU_NARROW_DRAM_BYTES_PER_PAIR_SITE = (6.072805e9 + 6.078759e9) / (37888 * 10000)  # profiles/r1_v4_decodeNarrow_s69_ncu_full.txt
NCU_DRAM_BYTES_PER_PAIR_SITE = (109.922461e9 + 109.458343e9) / (37888 * 10000)  # profiles/r1_v4_decodeFast_s69_ncu_full.txt
WORKLOAD = "cfg2: all-pairs, hashing off, 1000 haplotypes x 10000 SNPs, S=69 (30-100-2000), time=50, batchSize=32"


def dataset_root(rank):
    return f"/tmp/fsmc_bench/r{rank}/cfg2_{N_HAPS}x{N_SITES}"


def make_dataset(rank):
    from fastsmc_b200 import synth
    root = dataset_root(rank)
    if not os.path.exists(root + ".hap.gz"):
        synth.dataset(root, N_HAPS, N_SITES, SPAN_BP, CHROM, SEED)
    return root


def all_pairs_in_reference_order(n_ind):
    """hapA, hapB of HMM::decodeAll's enumeration (ref: ASMC_SRC/SRC/HMM.cpp:325-357), vectorised."""
    a, b = [], []
    for i in range(n_ind):
        j = np.repeat(np.arange(i), 4)
        ih = np.tile(np.array([0, 0, 1, 1]), i)
        jh = np.tile(np.array([0, 1, 0, 1]), i)
        a.append(2 * j + jh)
        b.append(np.full(4 * i, 2 * i) + ih)
        a.append(np.array([2 * i]))
        b.append(np.array([2 * i + 1]))
    return np.concatenate(a).astype(np.uint32), np.concatenate(b).astype(np.uint32)


def tiles_for(a, b, sites):
    n = len(a)
    T = (n + 31) // 32
    A = np.zeros(T * 32, np.uint32)
    B = np.zeros(T * 32, np.uint32)
    A[:n], B[:n] = a, b
    tp = np.full(T, 32, np.int32)
    tp[-1] = n - 32 * (T - 1)
    z, e = np.zeros(T, np.int32), np.full(T, sites, np.int32)
    return dict(hapA=A.reshape(T, 32), hapB=B.reshape(T, 32), tilePairs=tp, tileFrom=z, tileTo=e, tileScanFrom=z,
                tileScanTo=e, rows=np.arange(n))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        ok = [r for r in self.rows if len(r) == 7 and r[0].isdigit()]
        if not ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i] == "Active" for r in ok)]
        return {"sm_mhz": float(np.median([int(r[0]) for r in ok])), "sm_max_mhz": float(ok[0][1]),
                "power_w_max": max(float(r[2]) for r in ok), "samples": len(ok), "reasons": reasons}


REF_JOBS = 64  # the CPU arm cuts cfg2 into 64 jobs (reference jobs/jobInd) and runs one job per host core at a time


def reference_sample(cores):
    """One bounded sample of cfg2 on the reference's own CPU implementation: `cores` concurrent processes (the reference is
    single-threaded; its way to use many cores is one process per job, cpp_example/FastSMC_example_multiple_jobs.sh), each
    ASMC::FastSMC(params).run() of one of the REF_JOBS jobs of the data set, all 10,000 sites, same flags as the GPU arm.
    Returns (pair_sites, seconds of the slowest process's run(), kind, sample description).  Time excludes reading the
    data and preparing the model, as the GPU arm's does."""
    import re
    from oracle import pyoracle
    root = make_dataset(0)
    if pyoracle.reference_binary("avx") is None:
        return None
    procs = []
    # off-diagonal jobs only (job ids that are perfect squares are the diagonal jobs, a quarter of the pairs), so that the
    # concurrent processes carry equal work and the slowest one does not understate the CPU rate
    off_diagonal = [j for j in range(1, REF_JOBS + 1) if int(j ** 0.5) ** 2 != j]
    for k in range(cores):
        argv = pyoracle.reference_command("avx", root, DQ, f"/tmp/fsmc_bench/ref_out_{k}", hashing=0, jobs=REF_JOBS,
                                          jobInd=off_diagonal[k % len(off_diagonal)], time=50, noConditionalAgeEstimates=1,
                                          batchSize=32)
        procs.append(subprocess.Popen(argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    pairs, slowest = 0, 0.0
    for pr in procs:
        out, err = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("reference build failed: " + err[-300:])
        slowest = max(slowest, json.loads(out.strip().splitlines()[-1])["run_s"])
        pairs += int(re.search(r"Decoded (\d+) pairs", out + err).group(1))
    sample = (f"{cores} concurrent processes of the reference (unmodified sources, -O3 -DAVX -mavx), each one job of cfg2 cut "
              f"into {REF_JOBS} jobs (jobs/jobInd): {pairs} pairs x {N_SITES} sites per step")
    return float(pairs) * N_SITES, slowest, "reference", sample


def port_sample(cores, n_pairs):
    from oracle import pyoracle
    pyoracle.build()
    _declare_sample(pyoracle)
    root = make_dataset(0)
    o = pyoracle.Oracle(root, DQ, "/tmp/fsmc_bench/ref_out", hashing=False, time=50, noConditionalAgeEstimates=True,
                        doPerPairMAP=True, doPerPairPosteriorMean=True, batchSize=32)
    t0 = time.time()
    ps = pyoracle.lib().fo_decode_sample(o._h, n_pairs, cores)
    return ps, time.time() - t0, "port", f"first {n_pairs} pairs of the job x {N_SITES} sites, {cores} threads (oracle port)"


def run_reference(args, rank, world):
    """CPU arm: the reference's own implementation (oracle/_ref/ref_fastsmc_avx, kind "reference"; the oracle port if that
    binary is absent) on all host cores, each step a bounded sample of cfg2."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    times, pair_sites, kind, sample = [], 0.0, "port", ""
    for it in range(args.warmup + args.steps):
        r = reference_sample(cores) or port_sample(cores, 32 * max(cores * 2, 16))
        if it >= args.warmup:
            times.append(r[1])
            pair_sites, kind, sample = r[0], r[2], r[3]
    ms = 1e3 * float(np.mean(times))
    value = pair_sites / (ms / 1e3)
    line = {"impl": "reference", "metric": "hmm_pair_sites_per_s", "value": value, "unit": "pair-sites/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": "pair-sites/s", "cores": cores, "kind": kind, "sample": sample,
                             "per_thread": value / cores},
            "e2e": {"value": value, "unit": "pair-sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline():
    cores = os.cpu_count() or 1
    ps, dt, kind, sample = reference_sample(cores) or port_sample(cores, 64 * cores)
    return {"value": ps / dt, "unit": "pair-sites/s", "cores": cores, "kind": kind, "sample": sample + f", {dt:.1f} s",
            "per_thread": ps / dt / cores}


def max_over_ranks(value, world, device="cuda"):
    """Step time of the job = the slowest rank's (the only cross-rank operation of the benchmark)."""
    if world == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-run", action="store_true", help="skip the one-off FastSMC.run() wall-clock measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    from fastsmc_b200 import _native as N, asmc

    # ---- host layer: data set -> Data -> model tables (the product's own code; the oracle is not involved) ----------
    root = make_dataset(rank)
    p = asmc.DecodingParams()
    p.verbose = False
    p.inFileRoot, p.decodingQuantFile, p.outFileRoot = root, DQ, f"/tmp/fsmc_bench/r{rank}/out"
    p.decodingModeString, p.foldData, p.usingCSFS = "array", True, True
    p.FastSMC, p.hashing, p.batchSize, p.time = True, False, 32, 50
    p.noConditionalAgeEstimates = p.doPerPairMAP = p.doPerPairPosteriorMean = p.outputIbdSegmentLength = True
    p.useKnownSeed = True
    p.device = local
    p.validateParamsFastSMC()
    data = asmc.Data(p)
    tables = asmc.pyASMC.prepareModelTables(data, p)
    S, L = tables["emission1"].shape[1], data.sites

    ctx = N.Context(local)
    # a real (non-default) torch stream: handle 0 would mean "the context's own stream" to fsmc_ctx_set_stream
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_model(**tables)
    ctx.set_haplotypes(data.hapBits, L)
    a, b = all_pairs_in_reference_order(len(data.IIDList))
    tiles = tiles_for(a, b, L)
    flags = N.CALL_SEGMENTS | N.SEG_AGE
    pair_sites = float(len(a)) * L

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: pair list resident in HBM, kernels only ----------------------------------------------------------------
    plan = ctx.plan(tiles, flags, segment_capacity=1 << 21)
    for _ in range(args.warmup):
        plan.launch()
    torch.cuda.synchronize()
    r = plan.collect()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    with ClockSampler(local) as clocks:
        ev[0].record(stream)
        for k in range(args.steps):
            plan.launch()
            ev[k + 1].record(stream)
        barrier()
    per_step = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    r = plan.collect()
    n_segments = int(r.stats.numSegments)
    scratch_bytes = int(r.stats.scratchBytes)
    launches_per_step = int(r.stats.kernelLaunches)
    plan.close()
    total_ms = max_over_ranks(total_ms, world)
    ms_per_step = total_ms / args.steps
    value = world * pair_sites / (ms_per_step / 1e3)

    # ---- the same job with FastSMC's command-line default age estimates (conditional on TMRCA < time, i.e.
    # noConditionalAgeEstimates off: only the states below the threshold are consumed -> decodeNarrowKernel) ------------
    ctx2 = N.Context(local)
    ctx2.set_stream(stream.cuda_stream)
    ctx2.set_model(**dict(tables, age_threshold=tables["state_threshold"]))
    ctx2.set_haplotypes(data.hapBits, L)
    plan2 = ctx2.plan(tiles, flags, segment_capacity=1 << 21)
    for _ in range(args.warmup):
        plan2.launch()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        plan2.launch()
    e1.record(stream)
    barrier()
    narrow_ms = max_over_ranks(e0.elapsed_time(e1), world) / args.steps
    r2 = plan2.collect()
    narrow = {"value": world * pair_sites / (narrow_ms / 1e3), "unit": "pair-sites/s", "ms_per_step": narrow_ms,
              "kernel": "decodeNarrowKernel<69>" if r2.stats.narrowKernel else "decodeFastKernel<69>",
              "segments_per_step": int(r2.stats.numSegments), "scratch_bytes": int(r2.stats.scratchBytes),
              "flags": "as the headline run but age estimates conditional on TMRCA < time (FastSMC_exe default)"}
    plan2.close()
    ctx2.close()

    # ---- e2e: fsmc_decode with host buffers -----------------------------------------------------------------------------
    for _ in range(2):
        ctx.decode(tiles, flags, segment_capacity=1 << 21)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = ctx.decode(tiles, flags, segment_capacity=1 << 21)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    e2e_s = max_over_ranks(e2e_s, world)
    h2d = sum(tiles[k].nbytes for k in ("hapA", "hapB", "tilePairs", "tileFrom", "tileTo", "tileScanFrom", "tileScanTo"))
    d2h = int(res.stats.numSegments) * N.SEGMENT_DTYPE.itemsize + 16
    e2e = {"value": world * pair_sites / e2e_s, "unit": "pair-sites/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s,
           "call": "fsmc_decode (host pair lists in, host segment records out)"}

    # ---- one whole FastSMC.run(): read files, model preparation, decode, write .ibd.gz ------------------------------------
    ibd_wall = None
    if not args.no_e2e_run and rank == 0:
        t0 = time.perf_counter()
        f = asmc.FastSMC(p)
        f.run()
        ibd_wall = {"seconds": time.perf_counter() - t0, "segments": int(f.hmm().getNumberOfDetectedSegments()),
                    "decode_wall_s": f.hmm().getRunStats().decodeWallS, "output_wall_s": f.hmm().getRunStats().outputWallS,
                    "what": "FastSMC(params).run(): .hap.gz/.map/.samples + decoding quantities -> .ibd.gz"}
        del f

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        kernel_ms = float(np.mean(per_step))
        # algorithmic HBM bytes per pair-site of this kernel design: the backward sweep writes beta[S] floats and the
        # forward sweep reads them back (8*S), plus 2 bits of genotype input per pair-site (0.25 B)
        bytes_per_pair_site = 8.0 * S + 0.25
        achieved = pair_sites * bytes_per_pair_site / (kernel_ms / 1e3) / 1e9
        prop = torch.cuda.get_device_properties(local)
        fp32_peak = prop.multi_processor_count * 128 * 2 * 1.965e9 / 1e12
        fp32_achieved = pair_sites * FLOPS_PER_PAIR_SITE_STATE * S / (kernel_ms / 1e3) / 1e12
        line = {
            "metric": "hmm_pair_sites_per_s", "value": value, "unit": "pair-sites/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": int(len(a)), "sites": int(L), "states": int(S),
                       "pair_sites_per_step_per_gpu": pair_sites, "segments_per_step": n_segments,
                       "flags": "segment length + per-segment posterior mean + MAP over ALL states (noConditionalAgeEstimates), "
                                "the reference's regression-test / FastSMC-constructor defaults",
                       "l2": f"no flush needed: each step streams {scratch_bytes / 2**30:.0f} GiB of backward-sweep scratch "
                             "through HBM (>> 126 MB L2)", "parallelism": f"{world} independent jobs, one per GPU"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak,
                         # DRAM bytes per launch: dram__bytes_read.sum + dram__bytes_write.sum of the ncu --set full
                         # capture (profiles/r1_v4_decodeFast_s69_ncu_full.txt: 219.38 GB for 37 888 pairs x 10 000
                         # sites = 579.0 B per pair-site, the padding of 69 states to 72 included), scaled to this
                         # launch's pair-sites
                         "traffic": NCU_DRAM_BYTES_PER_PAIR_SITE * pair_sites,
                         "traffic_source": "ncu --set full capture of a 37888-pair launch, scaled by pair-sites",
                         "algorithmic_bytes": bytes_per_pair_site * pair_sites, "peak_source": peak_src,
                         "kernel": "decodeFastKernel<69>", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_pair_site": bytes_per_pair_site},
            "roofline_fp32": {"achieved": fp32_achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": fp32_achieved / fp32_peak,
                              "flops_per_pair_site": FLOPS_PER_PAIR_SITE_STATE * S,
                              "peak_source": "SMs x 128 lanes x 2 x 1.965 GHz (nominal)"},
            "default_flags": dict(narrow, roofline={"bound": "fp32", "achieved": narrow["value"] / world * FLOPS_PER_PAIR_SITE_STATE * S / 1e12,
                                                     "peak": fp32_peak, "unit": "TFLOP/s",
                                                     "frac": narrow["value"] / world * FLOPS_PER_PAIR_SITE_STATE * S / 1e12 / fp32_peak,
                                                     # DRAM bytes of one launch, from the ncu --set full capture
                                                     # profiles/r1_v4_decodeNarrow_s69_ncu_full.txt (12.15 GB for
                                                     # 37 888 pairs x 10 000 sites), scaled by pair-sites
                                                     "traffic": NCU_NARROW_DRAM_BYTES_PER_PAIR_SITE * pair_sites}),
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks.summary(),
            "ibd_wall": ibd_wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _declare_sample(pyoracle):
    import ctypes as C
    L = pyoracle.lib()
    L.fo_decode_sample.restype = C.c_double
    L.fo_decode_sample.argtypes = [C.c_void_p, C.c_long, C.c_int]


if __name__ == "__main__":
    main()
