#!/bin/bash
# full GPU pass: parity suite, smoke(), one bench line per BASELINE workload, reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_gl32_n1.json 2> gpurun_out/r2_bench_gl32_n1.err
python bench.py --workload kamani --steps 5 --warmup 3 > gpurun_out/r2_bench_kamani.json 2> gpurun_out/r2_bench_kamani.err
python bench.py --workload sine --steps 5 --warmup 3 > gpurun_out/r2_bench_sine.json 2> gpurun_out/r2_bench_sine.err
python bench.py --workload burgers --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_burgers.json 2> gpurun_out/r2_bench_burgers.err
python bench.py --workload gl_ref --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_gl_ref.json 2> gpurun_out/r2_bench_gl_ref.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
tail -16 gpurun_out/r2_pytest_gpu.log; tail -2 gpurun_out/r2_smoke.log
for f in gl32_n1 kamani sine burgers gl_ref reference_arm; do head -c 160 gpurun_out/r2_bench_$f.json; echo; done
