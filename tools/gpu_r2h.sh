#!/bin/bash
# GPU run r2h: nearest-code search with the block-per-query re-rank
mkdir -p gpurun_out
(python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "nearest" 2>&1 | grep -E "^\[|passed|failed|FAILED|^E  ") > gpurun_out/r2h_pytest.log 2>&1; tail -8 gpurun_out/r2h_pytest.log | cut -c1-220
python bench.py --config c5 > gpurun_out/r2h_bench_c5.json 2> gpurun_out/r2h_bench_c5.err; tail -2 gpurun_out/r2h_bench_c5.err; cut -c1-330 gpurun_out/r2h_bench_c5.json
python bench.py --config c5 --steps 1 --warmup 3 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 9 --csv --log-file gpurun_out/r2h_c5_launches.csv python bench.py --config c5 --steps 1 --warmup 3 > gpurun_out/r2h_ncu_c5.log 2>&1
grep -v "^==" gpurun_out/r2h_c5_launches.csv | cut -d, -f5,15 | tail -4
