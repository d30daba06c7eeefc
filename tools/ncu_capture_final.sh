#!/bin/bash
# Final ncu --set full pages of round 2 (run on the GPU box through gpurun; CSV pages land in gpurun_out/).
set -u
NCU="ncu --set full --clock-control none --kernel-name-base demangled"
# inference step: one launch of every kernel family of the final build (the 4th step of the run)
$NCU -k 'regex:conv_ln_hankel_persist|gemm_ln_tc_kernel|mlp_block_kernel|attention_packed_kernel|attention_tc_kernel|dwconv7_ln_w_kernel|layernorm_stream_kernel|tower_fwd_kernel|fusion_head_kernel|photo_embed_kernel' \
  --launch-skip 150 -c 50 -f -o gpurun_out/r2f_infer python bench.py --steps 1 --warmup 3 --blocks none --no-cpu-baseline > gpurun_out/ncu_r2f_infer.log 2>&1
ncu -i gpurun_out/r2f_infer.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_infer_final.csv 2>/dev/null; rm -f gpurun_out/r2f_infer.ncu-rep
# training step (eager launches so that every kernel is a separate node): the small-kernel families rewritten in the second half of the round
$NCU -k 'regex:layernorm_bwd_wide_kernel|attention_bwd_kernel|attention_packed_bwd_kernel|colsum_bf16x8_kernel|maxpool4_bwd_bf16x8_kernel|dropout_bf16x8_kernel|ew_bf16x8_kernel|act_bwd_bf16x8_kernel|layernorm_stream_kernel|adam_step_kernel|layernorm_bwd_reg_kernel' \
  --launch-skip 330 -c 120 -f -o gpurun_out/r2f_train python bench.py --workload train --graph off --steps 1 --warmup 3 > gpurun_out/ncu_r2f_train.log 2>&1
ncu -i gpurun_out/r2f_train.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_train_final.csv 2>/dev/null; rm -f gpurun_out/r2f_train.ncu-rep
# preprocessing kernels
$NCU -k 'regex:prep_lightcurve_kernel|prep_events_kernel|prep_spectrum_reg_kernel|cutout_median_reg_kernel|feature_sums_kernel' \
  --launch-skip 10 -c 10 -f -o gpurun_out/r2f_prep python bench.py --workload preprocess --prep-alerts 50000 > gpurun_out/ncu_r2f_prep.log 2>&1
ncu -i gpurun_out/r2f_prep.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_prep_final.csv 2>/dev/null; rm -f gpurun_out/r2f_prep.ncu-rep
wc -c gpurun_out/r2_ncu_full_*_final.csv
