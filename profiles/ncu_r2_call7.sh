set -x
D="python bench_extra.py ivf-q1 --rows 10000000 --profile-nq 4096 --iters 2"
$D > gpurun_out/ncu_plain10.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ivf_select_warp -s 1 -c 1 -o gpurun_out/prof_ivf_select_r2 $D > gpurun_out/ncu_l13.log 2>&1
echo rc=$?
