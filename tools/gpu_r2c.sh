#!/bin/bash
# GPU run r2c: full GPU suite after the backward-accumulator fix, gradient check, round-2 bench lines, ncu metric passes
mkdir -p gpurun_out
(python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[|   tap|passed|failed|FAILED|Error|^E  ") > gpurun_out/r2c_pytest.log 2>&1
tail -25 gpurun_out/r2c_pytest.log | cut -c1-260
GC="python tools/grad_check.py --config c2 --batch 8 --cache /tmp/gc_c2.pt"
($GC --precision fp32_parity; $GC --precision bf16; $GC --precision fp32_parity --steps 10; $GC --precision bf16 --steps 10) 2>&1 | grep -E "==|d loss|b256" > gpurun_out/r2c_gradcheck.log
cat gpurun_out/r2c_gradcheck.log | cut -c1-200
python bench.py --steps 5 --warmup 3 --layers-out gpurun_out/r2c_layers_c2.json > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err; tail -2 gpurun_out/r2c_bench_c2.err
python -c "import json; d=json.load(open('gpurun_out/r2c_bench_c2.json')); print(d['value'], d['e2e']['value'], d.get('fp32_parity'), d.get('gpu_reference'), d.get('cpu_baseline'))"
python bench.py --config c5 > gpurun_out/r2c_bench_c5.json 2> gpurun_out/r2c_bench_c5.err; tail -2 gpurun_out/r2c_bench_c5.err; cut -c1-330 gpurun_out/r2c_bench_c5.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err; cut -c1-600 gpurun_out/r2c_bench_reference.json
python bench.py --profile --steps 1 --warmup 3 > gpurun_out/r2c_profile_plain.json 2> gpurun_out/r2c_profile_plain.err &&
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.avg.per_cycle_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
    --clock-control none -k regex:tapgemm_kernel -s 819 -c 26 --csv --log-file gpurun_out/r2c_tapgemm_metrics.csv python bench.py --profile --steps 1 --warmup 3 > gpurun_out/r2c_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 300 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --profile --steps 1 --warmup 3 > gpurun_out/r2c_ncu2.log 2>&1
ls -la gpurun_out | grep r2c
