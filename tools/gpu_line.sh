#!/bin/bash
# line-marching Gauss-Seidel kernel: bit-identity against the one-launch-per-step kernel and timing per level
mkdir -p gpurun_out
{
B=32 NGRID=4 timeout 300 python tools/gs_bench.py 32 64 64 0 5
B=32 NGRID=3 timeout 300 python tools/gs_bench.py 32 32 32 2 5
B=32 NGRID=2 timeout 300 python tools/gs_bench.py 32 16 16 2 5
B=3 NGRID=2 timeout 300 python tools/gs_bench.py 8 16 16 0 5
D2=1 B=64 NGRID=6 DSF=1 timeout 300 python tools/gs_bench.py 256 256 0 5
D2=1 B=64 NGRID=5 DSF=1 timeout 300 python tools/gs_bench.py 128 128 0 5
} > gpurun_out/line_bench.log 2>&1
cat gpurun_out/line_bench.log
