#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
TAG=${ROUND_TAG:-r02zz}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${TAG}_smoke.log
timeout 200 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_n1.json
