#!/bin/bash
# 8-GPU run: config C3 (512x512, batch 128 split over the ranks, strong scaling), config C5 (2^20 codes sharded over 8 GPUs,
# NCCL all-gather + merge after the captured local search), and the in-process two-GPU test.
mkdir -p gpurun_out
# every command under its own timeout: a rank that hangs in teardown must not hold the box until gpurun's limit
TR="timeout 420 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --config c3 --steps 3 --warmup 3 > gpurun_out/r2_n8_bench_c3.json 2> gpurun_out/r2_n8_bench_c3.err; tail -2 gpurun_out/r2_n8_bench_c3.err; cut -c1-300 gpurun_out/r2_n8_bench_c3.json
$TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --config c5 --c5-total > gpurun_out/r2_n8_bench_c5.json 2> gpurun_out/r2_n8_bench_c5.err; tail -2 gpurun_out/r2_n8_bench_c5.err; cut -c1-400 gpurun_out/r2_n8_bench_c5.json
$TR --nproc-per-node 4 --master-port 29515 bench.py --gpus 4 --config c3 --steps 3 --warmup 3 > gpurun_out/r2_n4_bench_c3.json 2> gpurun_out/r2_n4_bench_c3.err; cut -c1-300 gpurun_out/r2_n4_bench_c3.json
$TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --config c5 --c5-total > gpurun_out/r2_n2_bench_c5.json 2> gpurun_out/r2_n2_bench_c5.err; cut -c1-300 gpurun_out/r2_n2_bench_c5.json
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "in_process_two_gpus" 2>&1 | tail -5) > gpurun_out/r2_n8_pytest_2gpu.log; cat gpurun_out/r2_n8_pytest_2gpu.log
