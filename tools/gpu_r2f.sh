#!/bin/bash
# GPU run r2f: nearest-code search after the re-rank rewrite (tests, bench, kernel list)
mkdir -p gpurun_out
(python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "nearest" 2>&1 | grep -E "^\[|passed|failed|FAILED|^E  ") > gpurun_out/r2f_pytest.log 2>&1; cat gpurun_out/r2f_pytest.log | cut -c1-220
python bench.py --config c5 > gpurun_out/r2f_bench_c5.json 2> gpurun_out/r2f_bench_c5.err; tail -2 gpurun_out/r2f_bench_c5.err; cut -c1-330 gpurun_out/r2f_bench_c5.json
python bench.py --config c5 --steps 1 --warmup 3 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 12 --csv --log-file gpurun_out/r2f_c5_launches.csv python bench.py --config c5 --steps 1 --warmup 3 > gpurun_out/r2f_ncu_c5.log 2>&1
grep -v "^==" gpurun_out/r2f_c5_launches.csv | cut -d, -f5,15 | tail -7
