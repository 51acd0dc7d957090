#!/bin/bash
for c in 1 0; do
  B=1 REPS=1 PDEOP_CHAIN_DEBUG=1 PDEOP_CHAIN=$c TAG=dbg_c$c timeout 300 python tools/cs_micro.py 2>&1 | grep -v Warn || echo "FAILED rc=$?"
done
python - <<PY
import numpy as np
a=np.load("gpurun_out/cs_out_dbg_c1.npy"); r=np.load("gpurun_out/cs_out_dbg_c0.npy")
print("fwd only: chain vs per-block: rel diff", np.linalg.norm(a-r)/np.linalg.norm(r))
# locate first differing entries in band order is not available (wave order); report by magnitude
d=np.abs(a-r); print("n", a.size, "num entries with rel diff >1e-6:", int((d>1e-6*np.abs(r).max()).sum()))
PY
