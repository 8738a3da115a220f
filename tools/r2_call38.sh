#!/bin/bash
TAG=${ROUND_TAG:-r02u}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
for i in 1 2; do
timeout 300 python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/${TAG}_c5_$i.json 2> gpurun_out/${TAG}_c5_$i.err
echo "c5 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_c5_$i.json'));print('%.1f img/s %.2f ms launches/step %d' % (d['value'], d['ms_per_step'], d['gpu_launches']/10))")"
done
