# AlignTransformer training step (forward + backward on the kernels): plain run first, then the ncu launch list
# and one full capture of the two attention-backward kernels.  Run on the GPU box through gpurun.
set -x
export PYTHONPATH=$PWD
NCU="ncu --clock-control none"
python bench.py --workload align_train --steps 5 > gpurun_out/r2_align_train.json 2> gpurun_out/r2_align_train.err || exit 1
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/r2_launches_align_train.csv python profiles/prof_align_train.py > gpurun_out/r2_ncu_align_train.log 2>&1
python profiles/prof_attn_bwd.py || exit 1
$NCU --set full --import-source on -k regex:"attn_bwd" -c 2 -o gpurun_out/r2_prof_attn_bwd -f python profiles/prof_attn_bwd.py > gpurun_out/r2_prof_attn_bwd.log 2>&1
