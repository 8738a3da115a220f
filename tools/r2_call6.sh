#!/bin/bash
# GPU call 6: full GPU test suite with the new defaults, then source-level ncu captures of the epilogue-bound GEMMs
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
R=${ROUND_TAG:-r02f}
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/${R}_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -n 12 gpurun_out/${R}_pytest_all.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${R}_smoke.log
for job in fc1_gelu_dg out_res dgrad_fc2_mul; do
  timeout 60 python tools/gemm_one.py $job > /dev/null 2>&1 && \
  timeout 150 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section LaunchStats \
    --clock-control none --import-source on -k regex:vitb_gemm -s 2 -c 1 -f -o "gpurun_out/gemm_${job}_${R}" \
    python tools/gemm_one.py $job > "gpurun_out/${R}_ncu_${job}.log" 2>&1; echo "ncu $job rc=$?"
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null
