for pair in 0 1; do for pol in 0 1 2; do
  timeout 120 python bench_extra.py batched --rows 10000000 --iters 10 --warmup 2 --tunable batch.cta_pair $pair --tunable batch.a_policy $pol 2>/dev/null | tail -1
done; done
