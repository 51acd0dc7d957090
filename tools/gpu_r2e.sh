#!/bin/bash
# round-2 evidence pass: bench line of the default workload, ncu launch list, ncu --set full captures of the top kernels
mkdir -p gpurun_out
timeout 300 python tools/diag_converged_time.py 2>&1 | grep -v Warn > gpurun_out/r2e_converged_time.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench_gl32.json 2> gpurun_out/r2e_bench_gl32.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/r2e_launches_gl32.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_ncu_launches.log 2>&1
for spec in "k_gs_cluster:20" "k_gs_fast:40" "k_apply:30" "k_band_chain:20" "k_syrk:200"; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/r2e_$k -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_ncu_$k.log 2>&1
done
cat gpurun_out/r2e_converged_time.log
head -c 400 gpurun_out/r2e_bench_gl32.json; echo
ls -la gpurun_out | grep r2e
