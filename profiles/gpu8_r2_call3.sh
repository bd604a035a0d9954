set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > gpurun_out/r2g_bench_1gpu.json 2> gpurun_out/r2g_bench_1gpu.err; echo bench1 rc=$?
for n in 2 4 8; do timeout 240 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2g_bench_${n}gpu.json 2> gpurun_out/r2g_bench_${n}gpu.err; echo bench$n rc=$?; done
timeout 300 $TR --nproc-per-node 8 --master-port 29541 tests/run_sharded_multi_gpu.py > gpurun_out/r2g_multi8.json 2> gpurun_out/r2g_multi8.err; echo multi_rc=$?
timeout 400 $TR --nproc-per-node 8 --master-port 29542 bench_extra.py sharded --rows 100000000 --data hier --nlist 16384 --recall-sweep 1 8 128 > gpurun_out/r2g_sharded_8gpu_100m.json 2> gpurun_out/r2g_sharded_8gpu_100m.err; echo sharded_rc=$?
