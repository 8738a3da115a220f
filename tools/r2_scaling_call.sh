#!/bin/bash
# Multi-GPU A/B of the data-parallel launch modes (run under `gpurun --gpus N`, N = 2 first, then 8):
#   graph1: whole step incl. ONE all-reduce as one CUDA graph | graph2: two graphs + eager all-reduce | overlap: eager, per-block
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash tools/r2_scaling_call.sh 2'
cd "$(dirname "$0")/.." || exit 1
N=${1:-2}
R=${ROUND_TAG:-r02}
mkdir -p gpurun_out
export VITB_BENCH_TIMEOUT_S=200
for mode in graph1 graph2 overlap; do
  timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29600 + N)) \
    bench.py --gpus "$N" --steps 20 --warmup 5 --no-cpu-baseline --ddp-mode $mode > "gpurun_out/${R}_scal_n${N}_${mode}.json" 2> "gpurun_out/${R}_scal_n${N}_${mode}.err"
  echo "n=$N $mode rc=$?"; cut -c1-200 "gpurun_out/${R}_scal_n${N}_${mode}.json"; tail -n 3 "gpurun_out/${R}_scal_n${N}_${mode}.err"
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > "gpurun_out/${R}_scal_n1.json" 2> /dev/null; echo "n=1 rc=$?"; cut -c1-200 "gpurun_out/${R}_scal_n1.json"
