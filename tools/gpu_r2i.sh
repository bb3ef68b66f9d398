#!/bin/bash
# GPU run r2i: ncu --set full of ONE rerank_kernel launch (details + per-line source counters)
mkdir -p gpurun_out
python bench.py --config c5 --steps 1 --warmup 3 > /dev/null 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:rerank_kernel -s 10 -c 1 -f -o /tmp/r2i_rerank python bench.py --config c5 --steps 1 --warmup 3 > gpurun_out/r2i_ncu.log 2>&1
ncu -i /tmp/r2i_rerank.ncu-rep --page details > gpurun_out/r2i_rerank_details.txt 2>&1
ncu -i /tmp/r2i_rerank.ncu-rep --page source --csv > gpurun_out/r2i_rerank_source.csv 2>&1
ls -la gpurun_out | grep r2i
grep -E "Duration|Executed Ipc|Warp Cycles Per Issued|Stall|stall|Theoretical Occ|Achieved Occ|Registers|L2 Hit|DRAM Throughput|Mem Busy|Max Bandwidth" gpurun_out/r2i_rerank_details.txt | head -40
