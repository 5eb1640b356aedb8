#!/bin/bash
# Round-2 measurement batch (one B200): bench lines, ncu launch list + full captures, cfg3 probe with and without the
# packed-matrix cache.  Everything lands in gpurun_out/.
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_final_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-jobs-run --no-e2e-run --no-cpu-baseline > gpurun_out/r2_final_bench_under_ncu.log 2>&1
FSMC_TRACE=1 python tools/scale_probe.py 10000 50000 240 1 1 1 gpurun_out/r2_final_scale_cfg3_reference_order.json > gpurun_out/r2_final_scale_cfg3.log 2>&1
FSMC_HAP_CACHE=1 python tools/scale_probe.py 10000 50000 240 1 1 1 gpurun_out/r2_final_scale_cfg3_cache_write.json > /dev/null 2>&1
FSMC_HAP_CACHE=1 python tools/scale_probe.py 10000 50000 240 1 1 1 gpurun_out/r2_final_scale_cfg3_cache_read.json > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:decodeNarrowKernel -c 1 -o gpurun_out/r2_final_narrowSparse -f \
    python bench.py --steps 1 --warmup 1 --no-jobs-run --no-e2e-run --no-cpu-baseline > gpurun_out/r2_final_ncu_sparse.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_final_narrowSparse.ncu-rep > gpurun_out/r2_final_decodeNarrowSparse_ncu_full.txt
ncu --set full --clock-control none --import-source on -k regex:refineKernel -c 1 -o gpurun_out/r2_final_refine -f \
    python bench.py --steps 1 --warmup 1 --no-jobs-run --no-e2e-run --no-cpu-baseline > gpurun_out/r2_final_ncu_refine.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_final_refine.ncu-rep > gpurun_out/r2_final_refine_ncu_full.txt
tail -c 600 gpurun_out/r2_final_bench.json
