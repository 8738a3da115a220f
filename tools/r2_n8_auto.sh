#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
N=${1:-8}
R=${ROUND_TAG:-r02y}
mkdir -p gpurun_out
export VITB_BENCH_TIMEOUT_S=150
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
  bench.py --gpus "$N" --steps 20 --warmup 5 --no-cpu-baseline > "gpurun_out/${R}_n${N}_auto.json" 2> "gpurun_out/${R}_n${N}_auto.err"
echo "n=$N auto rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${R}_n${N}_auto.json'));print('%.1f img/s %.2f ms loss %.5f | %s' % (d['value'], d['ms_per_step'], d['e2e']['last_loss'], d['config']['grad_exchange']))" 2>&1 | tail -1)"
tail -n 4 "gpurun_out/${R}_n${N}_auto.err" | cut -c1-200
