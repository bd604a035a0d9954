for t in "batch.cta_pair 0" "batch.cta_pair 1" "batch.growth 4" "batch.growth 6" "batch.first_chunk 4096" "batch.cap 6144"; do
  set -- $t
  timeout 120 python bench_extra.py batched --rows 10000000 --iters 10 --warmup 2 --tunable $1 $2 2>/dev/null | tail -1
done
