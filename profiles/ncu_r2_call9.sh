set -x
# launch list of the final build's bench command (N = 1 device stream = scan + finishing kernel per query)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
timeout 150 $B > gpurun_out/ncu_plain_final.log 2>&1 && timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ts:: -c 1000 --csv --log-file gpurun_out/launches_bench_r2_final.csv $B > gpurun_out/ncu_l_final.log 2>&1
echo launches_rc=$?
tail -2 gpurun_out/ncu_l_final.log | cut -c1-300
grep -c "ts::" gpurun_out/launches_bench_r2_final.csv
