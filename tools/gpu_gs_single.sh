#!/bin/bash
mkdir -p gpurun_out
{
for sg in 0 3 5; do
  echo "PDEOP_GS_SINGLE=$sg"
  PDEOP_GS_SINGLE=$sg B=32 NGRID=2 timeout 300 python tools/gs_bench.py 32 16 16 2 0 4
done
PDEOP_GS_SINGLE=10 B=32 NGRID=3 timeout 300 python tools/gs_bench.py 32 32 32 2 0
} 2>&1 | grep -v Warn > gpurun_out/gs_single.log
cat gpurun_out/gs_single.log
