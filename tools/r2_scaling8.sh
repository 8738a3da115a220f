#!/bin/bash
# 8-GPU call (billed 8x): the headline config in the single-graph data-parallel mode, the gradient exchange alone, and the
# two secondary configs BASELINE.json quotes at 8 GPUs.   gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_scaling8.sh'
cd "$(dirname "$0")/.." || exit 1
N=${1:-8}
R=${ROUND_TAG:-r02}
mkdir -p gpurun_out
export VITB_BENCH_TIMEOUT_S=170
run() {  # name, extra args
  local name=$1; shift
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus "$N" --steps 20 --warmup 5 --no-cpu-baseline "$@" > "gpurun_out/${R}_n${N}_${name}.json" 2> "gpurun_out/${R}_n${N}_${name}.err"
  echo "n=$N $name rc=$?"; cut -c1-220 "gpurun_out/${R}_n${N}_${name}.json"; tail -n 2 "gpurun_out/${R}_n${N}_${name}.err" | cut -c1-200
}
run c2_graph1 --config c2 --ddp-mode graph1
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29911 tools/allreduce_bench.py \
  > "gpurun_out/${R}_n${N}_allreduce.log" 2>&1; echo "allreduce rc=$?"; grep -i "all-reduce\|nvlink" "gpurun_out/${R}_n${N}_allreduce.log"
if [ "${SCALE_ALL:-1}" = "1" ]; then
  run c3_graph1 --config c3 --ddp-mode graph1
  run c5_graph1 --config c5 --ddp-mode graph1
fi
if [ "${SCALE_OVERLAP:-0}" = "1" ]; then run c2_overlap --config c2 --ddp-mode overlap; fi
