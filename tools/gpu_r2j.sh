#!/bin/bash
# GPU run r2j: CTA pairs in the nearest-code GEMM (A/B)
mkdir -p gpurun_out
(python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "nearest" 2>&1 | grep -E "^\[|passed|failed|FAILED|^E  ") > gpurun_out/r2j_pytest.log 2>&1; tail -8 gpurun_out/r2j_pytest.log | cut -c1-220
python bench.py --config c5 > gpurun_out/r2j_bench_c5_pair.json 2> gpurun_out/r2j_bench_c5_pair.err; tail -2 gpurun_out/r2j_bench_c5_pair.err; cut -c1-330 gpurun_out/r2j_bench_c5_pair.json
LA_NEAREST_PAIR=0 python bench.py --config c5 > gpurun_out/r2j_bench_c5_single.json 2> gpurun_out/r2j_bench_c5_single.err; cut -c1-330 gpurun_out/r2j_bench_c5_single.json
