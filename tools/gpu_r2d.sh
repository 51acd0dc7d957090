#!/bin/bash
mkdir -p gpurun_out
{
B=32 NGRID=4 timeout 300 python tools/gs_bench.py 32 64 64 0
PDEOP_VARIANT_SO=tools/libpdeop_noalloc.so B=32 NGRID=4 timeout 300 python tools/gs_bench.py 32 64 64 0
PDEOP_VARIANT_SO=tools/libpdeop_noalloc.so B=32 NGRID=3 timeout 300 python tools/gs_bench.py 32 32 32 2
PDEOP_VARIANT_SO=tools/libpdeop_noalloc.so D2=1 B=64 NGRID=6 DSF=1 timeout 300 python tools/gs_bench.py 256 256 0
} > gpurun_out/r2d_gs_noalloc.log 2>&1
grep -v Warn gpurun_out/r2d_gs_noalloc.log
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
tail -14 gpurun_out/r2d_pytest.log
