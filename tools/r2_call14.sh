#!/bin/bash
# round 2, call 14: fused class-token last block + single-query kernels + direct bias sums: tests, A/B bench, colsum sweep
TAG=${ROUND_TAG:-r02m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
for f in 0 1; do
  VITB_ROW0_FUSED=$f timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_row0_$f.json 2> gpurun_out/${TAG}_bench_row0_$f.err; echo "bench row0=$f rc=$?"; cut -c1-220 gpurun_out/${TAG}_bench_row0_$f.json
done
for f in 0 1; do
  VITB_ROW0_FUSED=$f timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_row0_${f}b.json 2> gpurun_out/${TAG}_bench_row0_${f}b.err; echo "bench row0=$f rc=$?"; cut -c1-220 gpurun_out/${TAG}_bench_row0_${f}b.json
done
timeout 200 python tools/colsum_bench.py > gpurun_out/${TAG}_colsum.log 2>&1; echo "colsum rc=$?"; cat gpurun_out/${TAG}_colsum.log
