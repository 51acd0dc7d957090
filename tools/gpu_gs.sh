#!/bin/bash
mkdir -p gpurun_out
{
B=32 NGRID=4 timeout 300 python tools/gs_bench.py 32 64 64 0 9 4
B=32 NGRID=3 timeout 300 python tools/gs_bench.py 32 32 32 0 9 4 10
D2=1 B=64 NGRID=6 DSF=1 timeout 300 python tools/gs_bench.py 256 256 0 9 10
D2=1 B=64 NGRID=4 DSF=1 timeout 300 python tools/gs_bench.py 64 64 0 9 10
D2=1 B=64 NGRID=3 DSF=1 timeout 300 python tools/gs_bench.py 32 32 0 9 10
} > gpurun_out/gs_bench.log 2>&1
cat gpurun_out/gs_bench.log
