set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29533 tests/run_sharded_multi_gpu.py > gpurun_out/r2_multi8.json 2> gpurun_out/r2_multi8.err; echo multi_rc=$?
for ex in pdl one-kernel nccl; do timeout 240 $TR --nproc-per-node 8 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 5 --exchange $ex > gpurun_out/r2_bench_8gpu_$ex.json 2> gpurun_out/r2_bench_8gpu_$ex.err; echo bench8_$ex rc=$?; done
timeout 240 $TR --nproc-per-node 8 --master-port 29535 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/r2_bench_8gpu_pdl_200.json 2> gpurun_out/r2_bench_8gpu_pdl_200.err; echo bench8_200 rc=$?
timeout 240 $TR --nproc-per-node 4 --master-port 29536 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_4gpu_pdl.json 2> gpurun_out/r2_bench_4gpu_pdl.err; echo bench4 rc=$?
timeout 400 $TR --nproc-per-node 8 --master-port 29537 bench_extra.py sharded --rows 100000000 --data hier --nlist 16384 --recall-sweep 1 8 128 > gpurun_out/r2_sharded_8gpu_100m.json 2> gpurun_out/r2_sharded_8gpu_100m.err; echo sharded_rc=$?
