# per-launch duration, tensor-pipe activity and DRAM traffic of the tcgen05 GEMM passes
# usage: bash profiles/quick_gemm.sh <tag>
tag=${1:-x}
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_active.avg
ncu --clock-control none --metrics $M -k regex:gemm_kernel -c 5 --csv --log-file gpurun_out/${tag}_c4_gemm.csv python bench.py --workload contrastive --steps 1 --warmup 1 --no-cpu > gpurun_out/${tag}_c4_gemm.log 2>&1
ncu --clock-control none --metrics $M -k regex:gemm_kernel -c 2 --csv --log-file gpurun_out/${tag}_c5_gemm.csv python bench.py --workload openvocab --steps 1 --warmup 1 --no-cpu --no-contrastive > gpurun_out/${tag}_c5_gemm.log 2>&1
python profiles/gemm_table.py gpurun_out/${tag}_c4_gemm.csv gpurun_out/${tag}_c5_gemm.csv
