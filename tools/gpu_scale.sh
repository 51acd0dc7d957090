#!/bin/bash
# multi-GPU measurements: BASELINE config 5 (GL 64x128x128, 32 instances per GPU) on N GPUs, and the two-device tests
mkdir -p gpurun_out
N=${1:-8}
python bench.py --gpus $N --workload gl64 --steps 1 --warmup 1 > gpurun_out/r2_bench_gl64_n$N.json 2> gpurun_out/r2_bench_gl64_n$N.err
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "two_devices or wrong_device" > gpurun_out/r2_pytest_2gpu.log 2>&1
head -c 600 gpurun_out/r2_bench_gl64_n$N.json; echo; tail -3 gpurun_out/r2_pytest_2gpu.log; tail -5 gpurun_out/r2_bench_gl64_n$N.err
