#!/bin/bash
# round 2, call 17: dropout kernels + module path, colsum / LayerNorm vector reductions: tests, bench
TAG=${ROUND_TAG:-r02n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
for i in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_$i.json 2> gpurun_out/${TAG}_bench_$i.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_$i.json
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
