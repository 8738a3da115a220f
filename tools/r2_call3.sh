#!/bin/bash
# third GPU call: first run of the persistent warp-specialised attention kernels (vitb_attention_ws.cu)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
R=${ROUND_TAG:-r02c}
timeout 120 python tools/attn_ws_check.py fwd > gpurun_out/${R}_ws_fwd.log 2>&1; echo "ws fwd rc=$?"; tail -n 14 gpurun_out/${R}_ws_fwd.log
timeout 120 python tools/attn_ws_check.py bwd > gpurun_out/${R}_ws_bwd.log 2>&1; echo "ws bwd rc=$?"; tail -n 14 gpurun_out/${R}_ws_bwd.log
timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attn_ws" > gpurun_out/${R}_pytest_ws.log 2>&1; echo "pytest ws rc=$?"; tail -n 3 gpurun_out/${R}_pytest_ws.log
timeout 200 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu > gpurun_out/${R}_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -n 3 gpurun_out/${R}_pytest_gemm.log
VITB_ATTN_WS=1 timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/${R}_pytest_all_ws.log 2>&1; echo "pytest all (ws) rc=$?"; tail -n 3 gpurun_out/${R}_pytest_all_ws.log
VITB_ATTN_WS=1 VITB_QKV_MERGE=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${R}_bench_ws_nomerge.json 2> gpurun_out/${R}_bench_ws_nomerge.err; echo "bench ws nomerge rc=$?"; cut -c1-200 gpurun_out/${R}_bench_ws_nomerge.json
VITB_ATTN_WS=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${R}_bench_ws.json 2> gpurun_out/${R}_bench_ws.err; echo "bench ws rc=$?"; cut -c1-200 gpurun_out/${R}_bench_ws.json
VITB_ATTN_WS=1 timeout 60 python tools/attn_one.py > /dev/null 2>&1 && \
VITB_ATTN_WS=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 4 -c 2 -f -o gpurun_out/attn_ws_${R} \
  python tools/attn_one.py > gpurun_out/${R}_ncu_attn_ws.log 2>&1; echo "ncu attn ws rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
