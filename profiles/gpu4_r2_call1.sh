set -x
# configs[3] / configs[4] at their named size on 4 GPUs (25M rows per GPU)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 170 $TR --nproc-per-node 4 --master-port 29551 bench_extra.py sharded --rows 100000000 --data hier --nlist 16384 > gpurun_out/r2n_sharded_4gpu_100m.json 2> gpurun_out/r2n_sharded_4gpu_100m.err; echo sharded_rc=$?
tail -c 1500 gpurun_out/r2n_sharded_4gpu_100m.json
tail -3 gpurun_out/r2n_sharded_4gpu_100m.err
