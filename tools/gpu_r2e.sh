#!/bin/bash
# GPU run r2e: kernel lists of the C5 search and of the four-term step; ncu --set full over the 26 tap-GEMM launches of one step
# (report kept on the box, only the raw-page CSV comes back)
mkdir -p gpurun_out
python bench.py --config c5 --steps 1 --warmup 3 > gpurun_out/r2e_bench_c5.json 2>/dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv --log-file gpurun_out/r2e_c5_launches.csv python bench.py --config c5 --steps 1 --warmup 3 > gpurun_out/r2e_ncu_c5.log 2>&1
grep -v "^==" gpurun_out/r2e_c5_launches.csv | cut -d, -f5,15 | tail -12
python bench.py --author-weights --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_c2_author.json 2> gpurun_out/r2e_bench_c2_author.err; tail -2 gpurun_out/r2e_bench_c2_author.err; cut -c1-200 gpurun_out/r2e_bench_c2_author.json
python bench.py --author-weights --profile --steps 1 --warmup 3 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 6500 -c 700 --csv --log-file gpurun_out/r2e_author_launches.csv python bench.py --author-weights --profile --steps 1 --warmup 3 > gpurun_out/r2e_ncu_author.log 2>&1
python bench.py --profile --steps 1 --warmup 3 > gpurun_out/r2e_profile_plain.json 2> gpurun_out/r2e_profile_plain.err &&
ncu --set full --clock-control none -k regex:tapgemm_kernel -s 819 -c 26 -f -o /tmp/r2e_tapgemm_full python bench.py --profile --steps 1 --warmup 3 > gpurun_out/r2e_ncu_full.log 2>&1
ncu -i /tmp/r2e_tapgemm_full.ncu-rep --page raw --csv > gpurun_out/r2e_tapgemm_full_raw.csv 2> gpurun_out/r2e_ncu_export.log
ls -la gpurun_out | grep r2e
