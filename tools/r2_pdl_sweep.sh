#!/bin/bash
# programmatic dependent launch per kernel family (VITB_PDL_EXPERIMENTAL bit mask: 1 GEMM, 2 LayerNorm, 4 attention)
TAG=${ROUND_TAG:-r02p}
mkdir -p gpurun_out
for m in 0 1 2 4 7 0 3; do
  VITB_PDL_EXPERIMENTAL=$m timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_pdl_$m.json 2> gpurun_out/${TAG}_pdl_$m.err
  echo "pdl mask $m rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_pdl_$m.json'));print('%.1f img/s %.3f ms' % (d['value'], d['ms_per_step']))")"
done
