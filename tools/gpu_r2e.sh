#!/bin/bash
# round-2 evidence pass: bench line of the default workload, ncu launch list, ncu --set full captures of the top kernels
# (raw/details pages exported as CSV on the box; only the Gauss-Seidel report itself is kept: gpurun_out/ is capped at 64 MiB)
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench_gl32.json 2> gpurun_out/r2g_bench_gl32.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/r2g_launches_gl32.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_ncu_launches.log 2>&1
cap() {   # name, kernel regex (demangled name), skip, count
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
    -o gpurun_out/r2g_$1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_ncu_$1.log 2>&1
  ncu -i gpurun_out/r2g_$1.ncu-rep --page raw --csv > gpurun_out/r2g_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2g_$1.ncu-rep --page details --csv > gpurun_out/r2g_$1_details.csv 2>/dev/null
}
cap gs_cluster "k_gs_cluster" 20 1
cap gs_fast "k_gs_fast" 40 2
rm -f gpurun_out/r2g_gs_fast.ncu-rep
cap apply "k_apply" 0 3
rm -f gpurun_out/r2g_apply.ncu-rep
cap band_chain "k_band_chain" 20 1
rm -f gpurun_out/r2g_band_chain.ncu-rep
cap syrk128 "k_syrk<.*128" 60 1
rm -f gpurun_out/r2g_syrk128.ncu-rep
head -c 300 gpurun_out/r2g_bench_gl32.json; echo
du -sh gpurun_out; ls -la gpurun_out
