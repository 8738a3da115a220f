#!/bin/bash
# GPU call 11: tests after the GEMM clean-up / bias-gradient change, PDL A/B, Res-ViT eval with token compaction (graphs),
# secondary configs on one GPU, DRAM traffic of the GEMM launches of a step
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
R=${ROUND_TAG:-r02k}
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/${R}_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -n 3 gpurun_out/${R}_pytest_all.log
for pdl in 0 1 0 1; do
  VITB_PDL=$pdl timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${R}_bench_pdl${pdl}.json 2> gpurun_out/${R}_bench_pdl${pdl}.err; echo "bench pdl=$pdl rc=$?"; cut -c1-160 gpurun_out/${R}_bench_pdl${pdl}.json
done
VITB_PDL=1 timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/${R}_pytest_pdl.log 2>&1; echo "pytest pdl rc=$?"; tail -n 2 gpurun_out/${R}_pytest_pdl.log
timeout 300 python tools/resvit_eval_bench.py 128 > gpurun_out/${R}_resvit_eval.log 2>&1; echo "resvit eval rc=$?"; tail -n 6 gpurun_out/${R}_resvit_eval.log
for c in c3 c4 c5; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_${c}.json 2> gpurun_out/${R}_bench_${c}.err; echo "bench $c rc=$?"; cut -c1-200 gpurun_out/${R}_bench_${c}.json
done
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_${R}.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${R}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
