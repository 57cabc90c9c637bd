# Round-2 evidence capture (run on the GPU box through gpurun, after the plain runs exited 0):
#   launch lists (per-kernel durations) of the default bench, the contrastive step and the preprocessing,
#   ncu --set full captures of the kernels that changed this round.
set -x
NCU="ncu --clock-control none"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_plain_default.json 2> gpurun_out/r2_plain_default.err || exit 1
$NCU --metrics gpu__time_duration.sum -k regex:"sim_small|merge_partials|prep_rows|upsample|gemm_kernel|z_finalize|mpnce|pp_|attn_kernel|pair_coef|sum_partials" -c 400 --csv --log-file gpurun_out/r2_launches_default.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_ncu_default.log 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:"gemm_kernel|mpnce|prep_rows|pair_coef|pick_scale|sum_partials|z_finalize" -c 45 --csv --log-file gpurun_out/r2_launches_c4.csv python bench.py --workload contrastive --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_ncu_c4.log 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:"pp_" -c 16 --csv --log-file gpurun_out/r2_launches_pp.csv python bench.py --workload preprocess --steps 3 --no-cpu > gpurun_out/r2_ncu_pp.log 2>&1
$NCU --set full --import-source on -k regex:"sim_small_kernel|merge_partials" -c 2 -o gpurun_out/r2_prof_cls python bench.py --steps 1 --warmup 3 --no-cpu --no-addons > gpurun_out/r2_prof_cls.log 2>&1
$NCU --set full --import-source on -k regex:"gemm_kernel|mpnce|prep_rows_bwd" -c 9 -o gpurun_out/r2_prof_c4 python bench.py --workload contrastive --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_prof_c4.log 2>&1
$NCU --set full --import-source on -k regex:"pp_" -c 4 -o gpurun_out/r2_prof_pp python bench.py --workload preprocess --steps 3 --no-cpu > gpurun_out/r2_prof_pp.log 2>&1
$NCU --set full --import-source on -k regex:"gemm_kernel|z_finalize" -c 3 -o gpurun_out/r2_prof_c5 python bench.py --workload openvocab --steps 1 --warmup 3 --no-cpu --no-addons > gpurun_out/r2_prof_c5.log 2>&1
ls -la gpurun_out | tail -12
