# Round-1 evidence for the widened path (bench.py --workload align): plain run first, then the launch list of
# the same command, then ONE `--set full` capture of the seven kernels of an AlignTransformer layer.
set -x
NCU="ncu --clock-control none"
python bench.py --workload align --steps 5 --warmup 3 --no-cpu > gpurun_out/r1_plain_align.json 2> gpurun_out/r1_plain_align.err || exit 1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r1_launches_align.csv python bench.py --workload align --steps 3 --warmup 3 --no-cpu > gpurun_out/r1_ncu_align.log 2>&1
# launches matching the regex per step: 5 prep_rows (4 LayerNorms + the prompts), 8 GEMMs, 2 attention = 15;
# skip the first step and take LN1, qkv, attention, proj, LN2, fc1, fc2 of the next one
$NCU --set full --import-source on -k regex:"attn_kernel|gemm_kernel|prep_rows_kernel" -s 15 -c 7 -o gpurun_out/r1_prof_align python bench.py --workload align --steps 1 --warmup 3 --no-cpu > gpurun_out/r1_prof_align.log 2>&1
ls -la gpurun_out | tail -8
