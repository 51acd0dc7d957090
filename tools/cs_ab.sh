#!/bin/bash
# coarse-solve A/B: persistent chain kernel (PDEOP_CHAIN=1) vs one launch per block row (0), results compared
for b in 4 32; do
  for c in 1 0; do
    echo "B=$b PDEOP_CHAIN=$c"
    B=$b PDEOP_CHAIN=$c TAG=b${b}_c$c timeout 300 python tools/cs_micro.py || echo "FAILED rc=$?"
  done
  python - <<PY
import numpy as np
a=np.load("gpurun_out/cs_out_b${b}_c1.npy"); r=np.load("gpurun_out/cs_out_b${b}_c0.npy")
print("B=$b chain vs per-block: rel diff", np.linalg.norm(a-r)/np.linalg.norm(r), "max abs", np.abs(a-r).max())
PY
done
rm -f gpurun_out/cs_out_*.npy
