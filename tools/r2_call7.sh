#!/bin/bash
# GPU call 7: attention A/B after the shared-O / fused-colsum changes, test suite, bench, ONE source-level GEMM capture
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
R=${ROUND_TAG:-r02g}
timeout 120 python tools/attn_ws_check.py both > gpurun_out/${R}_ws_check.log 2>&1; echo "ws check rc=$?"; tail -n 8 gpurun_out/${R}_ws_check.log
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/${R}_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -n 3 gpurun_out/${R}_pytest_all.log
timeout 300 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/${R}_bench_n1.json
timeout 60 python tools/gemm_one.py fc1_gelu_dg > /dev/null 2>&1 && \
timeout 150 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section LaunchStats \
    --clock-control none --import-source on -k regex:vitb_gemm -s 2 -c 1 -f -o "gpurun_out/gemm_fc1_gelu_dg_${R}" \
    python tools/gemm_one.py fc1_gelu_dg > "gpurun_out/${R}_ncu_fc1.log" 2>&1; echo "ncu fc1 rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
