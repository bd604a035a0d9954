set -x
timeout 400 python -m pytest tests/test_gpu_ivf.py tests/test_gpu_batched.py -m gpu -x -q > gpurun_out/r2_tests22.log 2>&1; echo tests_rc=$?; tail -3 gpurun_out/r2_tests22.log
D="python bench_extra.py ivf-q1 --rows 10000000 --profile-nq 4096 --iters 3"
$D > gpurun_out/ncu_plain11.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ts:: -c 600 --csv --log-file gpurun_out/launches_ivf_batch_r2e.csv $D > gpurun_out/ncu_l14.log 2>&1
echo ivf_launches_rc=$?
timeout 300 python bench_extra.py ivf --rows 40000000 --data hier --mma-modes 3 > gpurun_out/r2_ivf_hier40m_g.json 2> gpurun_out/r2_ivf_hier40m_g.err; echo ivf_rc=$?
