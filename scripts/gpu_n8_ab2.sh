#!/bin/bash
# second 8-GPU A/B: the v all-gather by peer stores from the normalise kernel with ONE CTA per SM, against NCCL's all-gather
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() {  # tag [ENV=VAL ...]
  tag=$1; shift
  env "$@" timeout 300 $TR --nproc-per-node $N --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-e2e --exchange peer \
      > gpurun_out/r02_bench_n${N}_k20_ab2_$tag.json 2> gpurun_out/r02_bench_n${N}_ab2_$tag.err
  echo "bench $tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_n${N}_k20_ab2_$tag.json").read())
    print("  $tag: value %.3f it/s  %.2f ms/step  exchange=%s" % (d["value"], d["ms_per_step"], d["exchange"]))
    print("  phases", d["phases_ms_per_step"])
except Exception as e:
    print("  $tag: no line", e)
PY
}
run hybrid HLV_PEER_ALLGATHER=nccl
run stores_unicast HLV_PEER_ALLGATHER=peer HLV_MULTICAST_STORE=0
run stores_multicast HLV_PEER_ALLGATHER=peer HLV_MULTICAST_STORE=1
run all_unicast HLV_PEER_ALLGATHER=peer HLV_MULTICAST=0 HLV_MULTICAST_STORE=0
