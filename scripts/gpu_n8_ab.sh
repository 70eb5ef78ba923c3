#!/bin/bash
# 8-GPU evidence run: rank-count invariance of T (2/4/8 ranks; NCCL and peer exchange), the exchange A/B on the headline bench,
# BASELINE config 4 (Pythia-1.4B, m=50, bf16 basis sharded over 8 GPUs).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -q -k rank_count > gpurun_out/r02_pytest_ranks_n$N.log 2>&1; echo "invariance rc=$?"; tail -3 gpurun_out/r02_pytest_ranks_n$N.log | cut -c1-300
run() {  # tag exchange [ENV=VAL ...]
  tag=$1; ex=$2; shift; shift
  env "$@" timeout 400 $TR --nproc-per-node $N --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --exchange $ex \
      > gpurun_out/r02_bench_n${N}_k20_$tag.json 2> gpurun_out/r02_bench_n${N}_$tag.err
  echo "bench $tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_n${N}_k20_$tag.json").read())
    print("  $tag: value %.3f it/s  %.2f ms/step  e2e %s  exchange=%s" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"],3), d["exchange"]))
    print("  phases", d["phases_ms_per_step"])
except Exception as e:
    print("  $tag: no line", e)
PY
}
run peer_multicast peer HLV_MULTICAST=1
run nccl nccl HLV_MULTICAST=0
run peer_unicast peer HLV_MULTICAST=0
run peer_rs_nccl_ag peer HLV_MULTICAST=1 HLV_PEER_ALLGATHER=nccl
timeout 900 $TR --nproc-per-node $N --master-port 29533 scripts/run_configs.py --config 4 > gpurun_out/r02_config4_n$N.json 2> gpurun_out/r02_config4_n$N.err; echo "config 4 rc=$?"
tail -c 1200 gpurun_out/r02_config4_n$N.json; echo
