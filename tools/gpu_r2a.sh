#!/bin/bash
# GPU run r2a: full GPU test suite + C2 / C3 / C5 bench lines (run under gpurun from the repo root)
mkdir -p gpurun_out
(python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[|rel_|passed|failed|Error|error|assert" | tail -150) > gpurun_out/r2a_pytest.log 2>&1
python bench.py --steps 5 --warmup 3 --layers-out gpurun_out/r2a_layers_c2.json > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
tail -5 gpurun_out/r2a_bench_c2.err
python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-fp32-parity > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
tail -3 gpurun_out/r2a_bench_c3.err
python bench.py --config c5 > gpurun_out/r2a_bench_c5.json 2> gpurun_out/r2a_bench_c5.err
tail -3 gpurun_out/r2a_bench_c5.err
cat gpurun_out/r2a_bench_c5.json
python bench.py --author-weights --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c2_author.json 2> gpurun_out/r2a_bench_c2_author.err
tail -3 gpurun_out/r2a_bench_c2_author.err
cat gpurun_out/r2a_bench_c2_author.json
tail -40 gpurun_out/r2a_pytest.log
