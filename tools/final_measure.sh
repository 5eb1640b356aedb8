#!/bin/bash
# Round-2 final measurement batch (one B200): bench lines, ncu launch list + full captures, cfg3 probe, racecheck of the
# new kernels.  Everything lands in gpurun_out/.
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_end_bench.json 2> gpurun_out/r2_end_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_end_bench_reference.json 2> gpurun_out/r2_end_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_end_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-jobs-run --no-e2e-run --no-cpu-baseline > gpurun_out/r2_end_bench_under_ncu.log 2>&1
FSMC_TRACE=1 python tools/scale_probe.py 10000 50000 240 1 1 1 gpurun_out/r2_end_scale_cfg3_reference_order.json > gpurun_out/r2_end_scale_cfg3.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/racecheck_probe.py > gpurun_out/r2_end_racecheck.log 2>&1
tail -5 gpurun_out/r2_end_racecheck.log
ncu --set full --clock-control none --import-source on -k regex:decodeLaneWideKernel -c 1 -o gpurun_out/r2_end_laneWide159 -f \
    python bench.py --steps 1 --warmup 1 --no-jobs-run --no-e2e-run --no-cpu-baseline > gpurun_out/r2_end_ncu_lanewide.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_end_laneWide159.ncu-rep > gpurun_out/r2_end_decodeLaneWide159_ncu_full.txt
ncu --set full --clock-control none --import-source on -k regex:decodeLaneKernel -c 1 -o gpurun_out/r2_end_lane159 -f \
    python bench.py --steps 1 --warmup 1 --no-jobs-run --no-e2e-run --no-cpu-baseline > gpurun_out/r2_end_ncu_lane.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_end_lane159.ncu-rep > gpurun_out/r2_end_decodeLane159_ncu_full.txt
tail -c 600 gpurun_out/r2_end_bench.json
