#!/bin/bash
# ncu capture of the line-marching Gauss-Seidel kernel (fine level, batch 32); the plain run comes first
mkdir -p gpurun_out
B=32 NGRID=4 REPS=2 timeout 300 python tools/gs_bench.py 32 64 64 5 > gpurun_out/line_ncu_plain.log 2>&1 || exit 1
B=32 NGRID=4 REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gs_line -s 1 -c 1 \
  -o gpurun_out/line_fine -f python tools/gs_bench.py 32 64 64 5 > gpurun_out/line_ncu.log 2>&1
ncu -i gpurun_out/line_fine.ncu-rep --page raw --csv > gpurun_out/r2_gs_line_raw.csv 2>/dev/null
ncu -i gpurun_out/line_fine.ncu-rep --page details --csv > gpurun_out/r2_gs_line_details.csv 2>/dev/null
rm -f gpurun_out/line_fine.ncu-rep
grep -v Warn gpurun_out/line_ncu_plain.log | tail -1
