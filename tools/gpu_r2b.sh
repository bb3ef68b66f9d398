#!/bin/bash
# GPU run r2b: locate the backward discrepancy at the C2 shape (per-layer gradient check under kernel switches),
# then the round-2 additions (1-pass nearest codes, micro-batches) and benches.
mkdir -p gpurun_out
L=gpurun_out/r2b_gradcheck.log
: > $L
GC="python tools/grad_check.py --config c2 --batch 8 --cache /tmp/gc_c2.pt"
$GC --precision fp32_parity >> $L 2>&1
$GC --precision bf16 >> $L 2>&1
$GC --precision fp32_parity --simt >> $L 2>&1
for sw in LA_CTA2=0 LA_CTA2=2 LA_BN=64 LA_NO_STAGED=1 LA_SEED_STREAM=1 LA_NO_FIR_TMA=1 LA_UPCONV_SPLIT_MIN_RES=100000; do
  env $sw $GC --precision fp32_parity 2>&1 | grep -E "==|d loss|b256|b128|Error|error" >> $L
done
python tools/grad_check.py --config c1 --batch 4 --cache /tmp/gc_c1.pt --precision fp32_parity 2>&1 | grep -E "==|d loss|b128|b64" >> $L
for st in 2 3 10; do
  $GC --precision fp32_parity --steps $st 2>&1 | grep "==" >> $L
  LA_NO_GRAPH=1 $GC --precision fp32_parity --steps $st 2>&1 | grep "==" >> $L
done
grep -v Warning $L | cut -c1-220
(python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "nearest or micro_batches" 2>&1 | grep -E "^\[|passed|failed|Error|assert" | tail -30) > gpurun_out/r2b_pytest.log 2>&1
cat gpurun_out/r2b_pytest.log | cut -c1-250
python bench.py --config c5 > gpurun_out/r2b_bench_c5.json 2> gpurun_out/r2b_bench_c5.err; tail -2 gpurun_out/r2b_bench_c5.err; cut -c1-400 gpurun_out/r2b_bench_c5.json
python bench.py --micro-batches 2 --steps 3 --no-cpu-baseline --no-gpu-reference --no-fp32-parity > gpurun_out/r2b_bench_c2_micro2.json 2> gpurun_out/r2b_bench_c2_micro2.err; tail -2 gpurun_out/r2b_bench_c2_micro2.err; cut -c1-300 gpurun_out/r2b_bench_c2_micro2.json
python bench.py --micro-batches 4 --steps 3 --no-cpu-baseline --no-gpu-reference --no-fp32-parity > gpurun_out/r2b_bench_c2_micro4.json 2> gpurun_out/r2b_bench_c2_micro4.err; cut -c1-300 gpurun_out/r2b_bench_c2_micro4.json
python bench.py --steps 3 --no-cpu-baseline --no-fp32-parity > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; tail -2 gpurun_out/r2b_bench_c2.err; python -c "import json; d=json.load(open('gpurun_out/r2b_bench_c2.json')); print(d['value'], d['e2e']['value'], d.get('gpu_reference'))"
