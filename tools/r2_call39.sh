#!/bin/bash
TAG=${ROUND_TAG:-r02v}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/${TAG}_c5.json 2> gpurun_out/${TAG}_c5.err
echo "c5 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_c5.json'));print('%.1f img/s %.2f ms' % (d['value'], d['ms_per_step']))")"
timeout 300 python bench.py > gpurun_out/${TAG}_c2.json 2> gpurun_out/${TAG}_c2.err
echo "c2 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_c2.json'));print('%.1f img/s %.2f ms e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))")"
