#!/bin/bash
# round-2 first GPU pass: parity suite, then one bench line per BASELINE workload
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --deselect "tests/test_gpu_parity.py::test_benchmarked_config_vs_oracle[2]" --durations=15 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_gl32.json 2> gpurun_out/r2a_bench_gl32.err
python bench.py --workload kamani --steps 5 --warmup 3 > gpurun_out/r2a_bench_kamani.json 2> gpurun_out/r2a_bench_kamani.err
python bench.py --workload sine --steps 5 --warmup 3 > gpurun_out/r2a_bench_sine.json 2> gpurun_out/r2a_bench_sine.err
python bench.py --workload burgers --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_burgers.json 2> gpurun_out/r2a_bench_burgers.err
tail -3 gpurun_out/r2a_pytest.log
head -c 600 gpurun_out/r2a_bench_gl32.json
