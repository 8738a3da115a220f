#!/bin/bash
# 2-GPU sanity run of the data-parallel step with the final binaries: c2 (one graph incl. the all-reduce), c5, and c2 in the
# eager overlapped-bucket mode (exercises the module hooks of the last block)
cd "$(dirname "$0")/.." || exit 1
N=2
R=${ROUND_TAG:-r02w}
mkdir -p gpurun_out
export VITB_BENCH_TIMEOUT_S=170
run() {
  local name=$1; shift
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus "$N" --steps 20 --warmup 5 --no-cpu-baseline "$@" > "gpurun_out/${R}_n${N}_${name}.json" 2> "gpurun_out/${R}_n${N}_${name}.err"
  echo "n=$N $name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${R}_n${N}_${name}.json'));print('%.1f img/s %.2f ms nvlink %s' % (d['value'], d['ms_per_step'], d.get('nvlink')))" 2>&1 | tail -1)"
}
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_n1_c2.json 2> gpurun_out/${R}_n1_c2.err
echo "n=1 c2 $(python -c "import json;d=json.load(open('gpurun_out/${R}_n1_c2.json'));print('%.1f img/s %.2f ms' % (d['value'], d['ms_per_step']))")"
run c2_graph1 --config c2 --ddp-mode graph1
run c5_graph1 --config c5 --ddp-mode graph1
run c2_overlap --config c2 --ddp-mode overlap
