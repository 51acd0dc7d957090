#!/bin/bash
# round-2 full GPU pass: parity suite, one bench line per BASELINE workload, reference arm, ncu launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r2c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench_gl32.json 2> gpurun_out/r2c_bench_gl32.err
python bench.py --workload kamani --steps 5 --warmup 3 > gpurun_out/r2c_bench_kamani.json 2> gpurun_out/r2c_bench_kamani.err
python bench.py --workload sine --steps 5 --warmup 3 > gpurun_out/r2c_bench_sine.json 2> gpurun_out/r2c_bench_sine.err
python bench.py --workload burgers --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_burgers.json 2> gpurun_out/r2c_bench_burgers.err
python bench.py --workload gl_ref --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_gl_ref.json 2> gpurun_out/r2c_bench_gl_ref.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2c_bench_reference_arm.json 2> gpurun_out/r2c_bench_reference_arm.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2c_launches_gl32.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu.log 2>&1
tail -25 gpurun_out/r2c_pytest.log
head -c 700 gpurun_out/r2c_bench_gl32.json
