#!/bin/bash
mkdir -p gpurun_out
for b in cublas cublaslt; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --prefill random --blas $b > gpurun_out/bench_blas_$b.json 2> gpurun_out/bench_blas_$b.err; echo "$b exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_blas_$b.json')); print('$b', round(d['value'],3), d['hvp_ms_per_step'], d['hvp_mode'][:12])"
done
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(round(d['value'],3), d['e2e']['value'], d['extras'])"
