#!/bin/bash
TAG=${ROUND_TAG:-r02q}
mkdir -p gpurun_out
for m in 5 4 0 5 4 0; do
  VITB_PDL_EXPERIMENTAL=$m timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_pdl_$m.json 2> gpurun_out/${TAG}_pdl_$m.err
  echo "pdl mask $m rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_pdl_$m.json'));print('%.1f img/s %.3f ms' % (d['value'], d['ms_per_step']))")"
done
VITB_PDL_EXPERIMENTAL=5 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_pdl5.log 2>&1; echo "pytest pdl=5 rc=$?"; tail -4 gpurun_out/${TAG}_pytest_pdl5.log
