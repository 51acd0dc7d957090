#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_gl32.json 2> gpurun_out/r2b_bench_gl32.err
python bench.py --workload kamani --steps 5 --warmup 3 > gpurun_out/r2b_bench_kamani.json 2> gpurun_out/r2b_bench_kamani.err
python bench.py --workload sine --steps 5 --warmup 3 > gpurun_out/r2b_bench_sine.json 2> gpurun_out/r2b_bench_sine.err
tail -30 gpurun_out/r2b_pytest.log
