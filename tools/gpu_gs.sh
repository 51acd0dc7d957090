#!/bin/bash
mkdir -p gpurun_out
{
B=32 NGRID=4 timeout 300 python tools/gs_bench.py 32 64 64 0
PDEOP_GS_THREADS=576 B=32 NGRID=4 timeout 300 python tools/gs_bench.py 32 64 64 0
PDEOP_GS_THREADS=640 B=32 NGRID=4 timeout 300 python tools/gs_bench.py 32 64 64 0
} > gpurun_out/gs_bench.log 2>&1
cat gpurun_out/gs_bench.log
