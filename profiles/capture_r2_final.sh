# Re-capture of the kernels that changed at the end of round 2 (small-N forward, preprocessing), same commands
# as capture_r2.sh; run on the GPU box through gpurun after the plain run exited 0.
set -x
NCU="ncu --clock-control none"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_plain_default.json 2> gpurun_out/r2_plain_default.err || exit 1
$NCU --metrics gpu__time_duration.sum -k regex:"sim_small|merge_partials|prep_rows|upsample|gemm_kernel|z_finalize|mpnce|pp_|attn_kernel|pair_coef|sum_partials" -c 400 --csv --log-file gpurun_out/r2_launches_default.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_ncu_default.log 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:"pp_" -c 16 --csv --log-file gpurun_out/r2_launches_pp.csv python bench.py --workload preprocess --steps 3 --no-cpu > gpurun_out/r2_ncu_pp.log 2>&1
$NCU --set full --import-source on -k regex:"pp_" -c 4 -o gpurun_out/r2_prof_pp -f python bench.py --workload preprocess --steps 3 --no-cpu > gpurun_out/r2_prof_pp.log 2>&1
$NCU --set full --import-source on -k regex:"sim_small_kernel|merge_partials" -c 2 -o gpurun_out/r2_prof_cls -f python bench.py --steps 1 --warmup 3 --no-cpu --no-addons > gpurun_out/r2_prof_cls.log 2>&1
tail -c 300 gpurun_out/r2_plain_default.json
