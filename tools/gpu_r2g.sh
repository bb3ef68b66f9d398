#!/bin/bash
# GPU run r2g: nearest-code search after the latency-oriented re-rank, discriminator / perceptual tests after their kernel changes
mkdir -p gpurun_out
(python -m pytest tests/test_gpu_parity.py tests/test_gpu_disc.py tests/test_gpu_lpips.py -q -m gpu -s -k "nearest or disc or lpips or realism or four_terms or plugin_runs" 2>&1 | grep -E "^\[|passed|failed|FAILED|^E  ") > gpurun_out/r2g_pytest.log 2>&1; tail -12 gpurun_out/r2g_pytest.log | cut -c1-220
python bench.py --config c5 > gpurun_out/r2g_bench_c5.json 2> gpurun_out/r2g_bench_c5.err; tail -2 gpurun_out/r2g_bench_c5.err; cut -c1-330 gpurun_out/r2g_bench_c5.json
python bench.py --author-weights --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_c2_author.json 2> gpurun_out/r2g_bench_c2_author.err; tail -2 gpurun_out/r2g_bench_c2_author.err; cut -c1-200 gpurun_out/r2g_bench_c2_author.json
python bench.py --config c5 --steps 1 --warmup 3 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 12 --csv --log-file gpurun_out/r2g_c5_launches.csv python bench.py --config c5 --steps 1 --warmup 3 > gpurun_out/r2g_ncu_c5.log 2>&1
grep -v "^==" gpurun_out/r2g_c5_launches.csv | cut -d, -f5,15 | tail -7
