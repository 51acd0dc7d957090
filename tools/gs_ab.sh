#!/bin/bash
# A/B timing of the Gauss-Seidel kernels at the three multigrid level sizes of the GL 32x64x64 workload
mkdir -p gpurun_out
for pipe in ${PIPES:-1 0}; do
  echo "PDEOP_GS_PIPE=$pipe"
  PDEOP_GS_PIPE=$pipe NGRID=4 python tools/gs_micro.py 32 64 64
  PDEOP_GS_PIPE=$pipe NGRID=3 python tools/gs_micro.py 32 32 32
  PDEOP_GS_PIPE=$pipe NGRID=2 python tools/gs_micro.py 32 16 16
done
