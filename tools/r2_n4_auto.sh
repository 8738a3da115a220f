#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
R=${ROUND_TAG:-r02zy}
mkdir -p gpurun_out
export VITB_BENCH_TIMEOUT_S=150
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
  bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > "gpurun_out/${R}_n4_auto.json" 2> "gpurun_out/${R}_n4_auto.err"
echo "n=4 auto rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${R}_n4_auto.json'));print('%.1f img/s %.2f ms | %s' % (d['value'], d['ms_per_step'], d['config']['grad_exchange']))" 2>&1 | tail -1)"
