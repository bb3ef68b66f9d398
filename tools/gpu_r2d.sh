#!/bin/bash
# GPU run r2d: graph-captured nearest-code search (C5) with its kernel list; ncu --set full over the 26 tap-GEMM launches of one step
mkdir -p gpurun_out
(python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "nearest or bn128" 2>&1 | grep -E "^\[|passed|failed|FAILED|^E  ") > gpurun_out/r2d_pytest.log 2>&1; cat gpurun_out/r2d_pytest.log | cut -c1-220
python bench.py --config c5 > gpurun_out/r2d_bench_c5.json 2> gpurun_out/r2d_bench_c5.err; tail -2 gpurun_out/r2d_bench_c5.err; cut -c1-330 gpurun_out/r2d_bench_c5.json
python bench.py --profile --steps 1 --warmup 3 > gpurun_out/r2d_profile_plain.json 2> gpurun_out/r2d_profile_plain.err &&
ncu --set full --clock-control none -k regex:tapgemm_kernel -s 819 -c 26 -f -o gpurun_out/r2d_tapgemm_full python bench.py --profile --steps 1 --warmup 3 > gpurun_out/r2d_ncu1.log 2>&1
python bench.py --config c5 --steps 1 --warmup 3 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r2d_c5_launches.csv python bench.py --config c5 --steps 1 --warmup 3 > gpurun_out/r2d_ncu2.log 2>&1
ls -la gpurun_out | grep r2d
