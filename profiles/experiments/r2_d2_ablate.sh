# pass D2 timing-only ablations (RZ_EXP_D2 bits; see rz_sim_bwd.cu): where does the pass lose its time?
for V in base d2x1 d2x3 d2x4 d2x7 d2x15; do
  if [ $V = base ]; then unset RZ_B200_LIB; else export RZ_B200_LIB=$PWD/radzero_b200/_build/variants/librz_$V.so; fi
  ncu --clock-control none --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    -k regex:"gemm_kernel" -c 10 --csv --log-file gpurun_out/d2abl_$V.csv python bench.py --workload contrastive --steps 1 --warmup 1 --no-cpu --no-addons > gpurun_out/d2abl_$V.log 2>&1
done
