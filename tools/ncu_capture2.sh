set -u
NCU="ncu --set full --clock-control none --kernel-name-base demangled"
$NCU -k 'regex:gemm_tc_kernel' --launch-skip 180 -c 60 -f -o gpurun_out/r2_full_gemm python bench.py --steps 1 --warmup 3 --blocks none --no-cpu-baseline > gpurun_out/ncu_full_gemm.log 2>&1
ncu -i gpurun_out/r2_full_gemm.ncu-rep --page raw --csv > gpurun_out/r2_full_gemm.csv 2>/dev/null; rm -f gpurun_out/r2_full_gemm.ncu-rep
$NCU -k 'regex:prep_events_kernel|feature_sums_kernel|prep_spectrum_kernel|cutout_median_kernel' --launch-skip 8 -c 8 -f -o gpurun_out/r2_full_prep2 python bench.py --workload preprocess --prep-alerts 50000 > gpurun_out/ncu_full_prep2.log 2>&1
ncu -i gpurun_out/r2_full_prep2.ncu-rep --page raw --csv > gpurun_out/r2_full_prep2.csv 2>/dev/null; rm -f gpurun_out/r2_full_prep2.ncu-rep
wc -c gpurun_out/r2_full_gemm.csv gpurun_out/r2_full_prep2.csv
