set -x
E="python bench_extra.py batched --rows 10000000 --iters 1 --warmup 0 --tunable batch.cta_pair 1"
$E > gpurun_out/ncu_plain9.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:batched_gemm_topk -s 7 -c 1 -o gpurun_out/prof_k3_pair_r2 $E > gpurun_out/ncu_l12.log 2>&1
echo k3pair_rc=$?
