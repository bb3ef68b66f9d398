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
