#!/bin/bash
# Multi-GPU A/B of the data-parallel launch modes (run under `gpurun --gpus N`, N = 2 first, then 8):
#   eager overlapped wrapper (ddp.DataParallel)  vs  fwd+bwd graph + one all-reduce + optimizer graph (--graph-ddp)
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash tools/r2_scaling_call.sh 2'
cd "$(dirname "$0")/.." || exit 1
N=${1:-2}
mkdir -p gpurun_out
for mode in eager graph; do
  flag=""; [ "$mode" = graph ] && flag="--graph-ddp"
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29600 + N)) \
    bench.py --gpus "$N" --steps 20 --warmup 5 --no-cpu-baseline $flag > "gpurun_out/scal_n${N}_${mode}.json" 2> "gpurun_out/scal_n${N}_${mode}.err"
  echo "n=$N $mode rc=$?"; cut -c1-220 "gpurun_out/scal_n${N}_${mode}.json"
done
