#!/bin/bash
# GPU call 9: deep-store GEMM configuration (3 stages, 4 TMA-store tiles per epilogue warp) — tests, per-shape A/B, step A/B
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
R=${ROUND_TAG:-r02i}
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu > gpurun_out/${R}_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -n 3 gpurun_out/${R}_pytest_gemm.log
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/${R}_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -n 3 gpurun_out/${R}_pytest_all.log
for ew in 0 1; do
  VITB_GEMM_DEEPSTORE=$ew timeout 120 python tools/gemm_bench.py > gpurun_out/${R}_gemm_bench_ds$ew.log 2>&1; echo "gemm_bench deepstore=$ew rc=$?"; cat gpurun_out/${R}_gemm_bench_ds$ew.log
  VITB_GEMM_DEEPSTORE=$ew timeout 120 python tools/epi_ab.py > gpurun_out/${R}_epi_ab_ds$ew.log 2>&1; head -n 8 gpurun_out/${R}_epi_ab_ds$ew.log
done
for ew in 0 1 0 1; do
  VITB_GEMM_DEEPSTORE=$ew timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${R}_bench_ds${ew}.json 2> gpurun_out/${R}_bench_ds${ew}.err; echo "bench deepstore=$ew rc=$?"; cut -c1-160 gpurun_out/${R}_bench_ds${ew}.json
done
timeout 120 python tools/attn_ws_check.py both 2>&1 | tail -n 4
timeout 200 python tools/resvit_eval_bench.py 128 > gpurun_out/${R}_resvit_eval.log 2>&1; echo "resvit eval rc=$?"; cat gpurun_out/${R}_resvit_eval.log | tail -n 6
