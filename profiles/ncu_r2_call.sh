set -x
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --ivf-rows 10000000"
$B > gpurun_out/ncu_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench_r2.csv $B > gpurun_out/ncu_l1.log 2>&1
echo launches_rc=$?
C="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
$C > gpurun_out/ncu_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 4 -c 1 -o gpurun_out/prof_scan_topk_r2 $C > gpurun_out/ncu_l2.log 2>&1
echo k2_rc=$?
ncu --set full --clock-control none --import-source on -k regex:normalize_cast -s 10 -c 1 -o gpurun_out/prof_k1_r2 $C > gpurun_out/ncu_l3.log 2>&1
echo k1_rc=$?
D="python bench_extra.py ivf-q1 --rows 10000000 --profile-nq 4096 --iters 2"
$D > gpurun_out/ncu_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ivf_grouped_umma -s 1 -c 1 -o gpurun_out/prof_k4d_umma_r2 $D > gpurun_out/ncu_l4.log 2>&1
echo k4d_rc=$?
E="python bench_extra.py batched --rows 10000000 --iters 1 --warmup 0"
$E > gpurun_out/ncu_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:batched_gemm_topk -s 7 -c 1 -o gpurun_out/prof_k3_r2 $E > gpurun_out/ncu_l5.log 2>&1
echo k3_rc=$?
ls -la gpurun_out/*.ncu-rep
