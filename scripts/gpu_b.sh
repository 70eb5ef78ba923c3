#!/bin/bash
# second GPU pass: full parity suite, headline bench (both arms), HVP kernel breakdown, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "exit $?" >> gpurun_out/bench_n1.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "exit $?" >> gpurun_out/bench_ref_n1.err
timeout 300 python scripts/hvp_profile.py 8 > gpurun_out/hvp_profile.log 2>&1
CMD="python bench.py --steps 2 --warmup 1 --prefill random --no-e2e --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/ncu_launches.log
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"cgs_|lanczos_update|normalize_store|multi_tensor" -c 12 -o gpurun_out/prof_r01 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/ncu_full.log
tail -3 gpurun_out/pytest_gpu.log gpurun_out/bench_n1.err gpurun_out/bench_ref_n1.err gpurun_out/ncu_launches.log gpurun_out/ncu_full.log
cat gpurun_out/bench_n1.json gpurun_out/bench_ref_n1.json
