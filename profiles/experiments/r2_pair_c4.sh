for X in 1 7; do
RZ_EXP_PAIR=$X ncu --clock-control none --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"gemm_kernel" -c 10 --csv --log-file gpurun_out/pair_$X.csv python bench.py --workload contrastive --steps 1 --warmup 3 --no-cpu --no-addons > gpurun_out/pair_$X.log 2>&1
done
RZ_EXP_PAIR=1 python bench.py --workload contrastive --steps 10 --warmup 3 --no-cpu --no-addons 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pairD2', d['ms_per_step'], d['loss'], d['grad_checksum'])"
RZ_EXP_PAIR=7 python bench.py --workload contrastive --steps 10 --warmup 3 --no-cpu --no-addons 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pairAll', d['ms_per_step'], d['loss'], d['grad_checksum'])"
python bench.py --workload contrastive --steps 10 --warmup 3 --no-cpu --no-addons 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('single', d['ms_per_step'], d['loss'], d['grad_checksum'])"
