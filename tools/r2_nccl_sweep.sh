#!/bin/bash
# 8-GPU call: which NCCL algorithm carries the 343.5 MB gradient all-reduce on this box, and do the knobs change it?
cd "$(dirname "$0")/.." || exit 1
N=${1:-8}
R=${ROUND_TAG:-r02t}
mkdir -p gpurun_out
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29300 + RANDOM % 300)) \
    tools/allreduce_bench.py > "gpurun_out/${R}_nccl_${name}.log" 2>&1
  echo "== $name rc=$?"; grep -i "all-reduce ViT-B\|all-reduce ViT-L" "gpurun_out/${R}_nccl_${name}.log"
}
run default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING
grep -i "nvls\|algo\|channels\|NCCL version" "gpurun_out/${R}_nccl_default.log" | sort | uniq -c | sort -rn | head -12 | cut -c1-220
run nvls NCCL_ALGO=NVLS
run ring NCCL_ALGO=Ring
run ring32 NCCL_ALGO=Ring NCCL_MIN_NCHANNELS=32
run tree NCCL_ALGO=Tree
run nvlstree NCCL_ALGO=NVLSTree
