set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests10.log 2>&1; echo tests_rc=$?; tail -3 gpurun_out/r2_tests10.log
D="python bench_extra.py ivf-q1 --rows 10000000 --profile-nq 4096 --iters 3"
$D > gpurun_out/ncu_plain6.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ts:: -c 600 --csv --log-file gpurun_out/launches_ivf_batch_r2b.csv $D > gpurun_out/ncu_l8.log 2>&1
echo ivf_launches_rc=$?
ncu --set full --clock-control none --import-source on -k regex:ivf_grouped_umma -s 1 -c 1 -o gpurun_out/prof_k4d_umma_r2c $D > gpurun_out/ncu_l9.log 2>&1
echo k4d_rc=$?
timeout 300 python bench_extra.py ivf --rows 40000000 --data hier --mma-modes 1 3 > gpurun_out/r2_ivf_hier40m_c.json 2> gpurun_out/r2_ivf_hier40m_c.err; echo ivf_rc=$?
