set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > gpurun_out/r2m_bench_1gpu.json 2> gpurun_out/r2m_bench_1gpu.err; echo bench1 rc=$?
for n in 2 4 8; do timeout 240 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2m_bench_${n}gpu.json 2> gpurun_out/r2m_bench_${n}gpu.err; echo bench$n rc=$?; done
