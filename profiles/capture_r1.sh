# Round-1 evidence capture (run on the GPU box through gpurun, after the plain runs exited 0):
#   launch lists (per-kernel durations) of the default bench and the other workloads,
#   ncu --set full captures of the dominant kernels, one per workload.
set -x
NCU="ncu --clock-control none"
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r1_plain_default.json 2> gpurun_out/r1_plain_default.err || exit 1
$NCU --metrics gpu__time_duration.sum -c 800 --csv --log-file gpurun_out/r1_launches_default.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r1_ncu_default.log 2>&1
for w in seg openvocab; do
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/r1_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-contrastive > gpurun_out/r1_ncu_$w.log 2>&1
done
$NCU --set full --import-source on -k regex:"sim_small_kernel|merge_partials|prep_rows_kernel" -c 3 -o gpurun_out/r1_prof_cls python bench.py --steps 1 --warmup 1 --no-cpu --no-contrastive > gpurun_out/r1_prof_cls.log 2>&1
$NCU --set full --import-source on -k regex:"upsample_kernel|sim_small_kernel" -c 2 -o gpurun_out/r1_prof_seg python bench.py --workload seg --steps 1 --warmup 1 --no-cpu --no-contrastive > gpurun_out/r1_prof_seg.log 2>&1
$NCU --set full --import-source on -k regex:"gemm_kernel|z_finalize" -c 3 -o gpurun_out/r1_prof_c5 python bench.py --workload openvocab --steps 1 --warmup 1 --no-cpu --no-contrastive > gpurun_out/r1_prof_c5.log 2>&1
$NCU --set full --import-source on -k regex:"gemm_kernel|mpnce|prep_rows_bwd" -c 14 -o gpurun_out/r1_prof_c4 python bench.py --workload contrastive --steps 1 --warmup 1 --no-cpu > gpurun_out/r1_prof_c4.log 2>&1
ls -la gpurun_out | tail -20
