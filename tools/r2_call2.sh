#!/bin/bash
# second GPU call of round 2: re-collect what call 1 lost (its gpurun_out exceeded 64 MiB): test summary, bench line,
# epilogue A/B (incl. VITB_EPI_ROWRES), step bench with ROWRES on, launch list, small ncu raw-metric CSVs.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
R=${ROUND_TAG:-r02b}
VITB_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/${R}_pytest.log
timeout 300 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/${R}_bench_n1.json
VITB_EPI_ROWRES=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${R}_bench_n1_rowres.json 2> gpurun_out/${R}_bench_n1_rowres.err; echo "bench rowres rc=$?"; cut -c1-200 gpurun_out/${R}_bench_n1_rowres.json
timeout 120 python tools/epi_ab.py > gpurun_out/${R}_epi_ab.log 2>&1; echo "epi_ab rc=$?"; cat gpurun_out/${R}_epi_ab.log
timeout 200 python tools/epi_ablate.py > gpurun_out/${R}_epi_ablate.log 2>&1; echo "epi_ablate rc=$?"
timeout 120 python tools/gemm_bench.py > gpurun_out/${R}_gemm_bench.log 2>&1; echo "gemm_bench rc=$?"; cat gpurun_out/${R}_gemm_bench.log
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${R}_bench_nograph.json 2> /dev/null && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${R}.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${R}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
