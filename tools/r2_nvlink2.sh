#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
N=${1:-2}
R=${ROUND_TAG:-r02x}
mkdir -p gpurun_out
export VITB_BENCH_TIMEOUT_S=170
for mode in nvlink graph1; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus "$N" --steps 20 --warmup 5 --no-cpu-baseline --ddp-mode $mode > "gpurun_out/${R}_n${N}_${mode}.json" 2> "gpurun_out/${R}_n${N}_${mode}.err"
  echo "n=$N $mode rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${R}_n${N}_${mode}.json'));print('%.1f img/s %.2f ms loss %.5f | %s' % (d['value'], d['ms_per_step'], d['e2e']['last_loss'], d['config']['grad_exchange']))" 2>&1 | tail -1)"
  tail -n 3 "gpurun_out/${R}_n${N}_${mode}.err" | cut -c1-200
done
