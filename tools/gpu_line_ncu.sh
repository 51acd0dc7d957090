#!/bin/bash
# ncu capture of the line-marching Gauss-Seidel kernel; the plain run comes first
mkdir -p gpurun_out
B=32 NGRID=4 REPS=2 timeout 300 python tools/gs_bench.py 32 64 64 5 > gpurun_out/line_ncu_plain.log 2>&1 || exit 1
B=32 NGRID=4 REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gs_line -s 1 -c 1 \
  -o gpurun_out/line_fine -f python tools/gs_bench.py 32 64 64 5 > gpurun_out/line_ncu.log 2>&1
tail -3 gpurun_out/line_ncu.log
