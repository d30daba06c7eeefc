#!/bin/bash
# Round-2 ncu --set full captures (run on the GPU box through gpurun; reports land in gpurun_out/).
#   bash tools/ncu_capture.sh infer|train|prep
set -u
NCU="ncu --set full --clock-control none --kernel-name-base demangled"
case "$1" in
  infer)
    $NCU -k 'regex:conv_ln_hankel_persist|gemm_tc_kernel<128, 2, 2>|layernorm_stream_kernel|dwconv7_ln_w_kernel|attention_packed_kernel|attention_tc_kernel|gemm_ln_tc_kernel|mlp_block_kernel' \
      --launch-skip 114 -c 76 -f -o gpurun_out/r2_full_infer python bench.py --steps 1 --warmup 3 --blocks none --no-cpu-baseline > gpurun_out/ncu_full_infer.log 2>&1
    ;;
  train)
    $NCU -k 'regex:attention_packed_bwd_kernel|attention_bwd_kernel|adam_step_kernel|wgrad_tc_kernel<256, 3>|layernorm_bwd_reg_kernel|dwconv7_wgrad' \
      --launch-skip 360 -c 130 -f -o gpurun_out/r2_full_train python bench.py --workload train --graph off --steps 1 --warmup 3 > gpurun_out/ncu_full_train.log 2>&1
    ;;
  prep)
    $NCU -k 'regex:prep_lightcurve_kernel|prep_events_kernel|prep_spectrum_kernel|cutout_median_kernel|feature_sums_kernel' \
      -c 10 -f -o gpurun_out/r2_full_prep python bench.py --workload preprocess --prep-alerts 50000 > gpurun_out/ncu_full_prep.log 2>&1
    ;;
esac
# the reports are large (the merge back is capped at 64 MiB): keep the raw-metric CSV page, drop the report
ncu -i gpurun_out/r2_full_$1.ncu-rep --page raw --csv > gpurun_out/r2_full_$1.csv 2>/dev/null
rm -f gpurun_out/r2_full_$1.ncu-rep
tail -c 300 gpurun_out/ncu_full_$1.log; wc -c gpurun_out/r2_full_$1.csv
