#!/bin/bash
# First GPU call of a new round (one B200, ~6-8 GPU-minutes): validates the tree, measures the step, breaks the
# epilogue-bound GEMMs down by ablation, A/Bs the experimental epilogue, and captures the ncu evidence the third
# session of round 1 could not (it ran out of GPU budget).  Every leg writes its own file under gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r2_first_call.sh'
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
R=${ROUND_TAG:-r02a}
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/${R}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/${R}_smoke.log
timeout 300 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/${R}_bench_n1.json
timeout 120 python tools/epi_ab.py > gpurun_out/${R}_epi_ab.log 2>&1; echo "epi_ab rc=$?"; tail -n 20 gpurun_out/${R}_epi_ab.log
timeout 200 python tools/epi_ablate.py > gpurun_out/${R}_epi_ablate.log 2>&1; echo "epi_ablate rc=$?"; tail -n 8 gpurun_out/${R}_epi_ablate.log
timeout 120 python tools/gemm_bench.py > gpurun_out/${R}_gemm_bench.log 2>&1; echo "gemm_bench rc=$?"
timeout 60 python tools/attn_bench.py > gpurun_out/${R}_attn_bench.log 2>&1; echo "attn_bench rc=$?"; tail -n 4 gpurun_out/${R}_attn_bench.log
# ncu: the launch list of a step (same command exited 0 just above, modulo the launch mode), then full captures
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${R}_bench_nograph.json 2> /dev/null && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${R}.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${R}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
for job in "gemm_fc1_gelu_dg:tools/gemm_one.py:fc1_gelu_dg:regex:vitb_gemm" "gemm_dgrad_fc2_mul:tools/gemm_one.py:dgrad_fc2_mul:regex:vitb_gemm" \
           "gemm_out_res:tools/gemm_one.py:out_res:regex:vitb_gemm" "attn:tools/attn_one.py::regex:attn_" \
           "ln_bwd:tools/ln_one.py::regex:ln_bwd"; do
  IFS=: read -r name script arg kflag kpat <<< "$job"
  timeout 60 python "$script" $arg > /dev/null 2>&1 && \
  timeout 150 ncu --set full --clock-control none --import-source on -k "$kflag:$kpat" -s 2 -c 2 -f -o "gpurun_out/${name}_${R}" \
    python "$script" $arg > "gpurun_out/${R}_ncu_${name}.log" 2>&1; echo "ncu $name rc=$?"
done
ls -la gpurun_out/*_${R}.ncu-rep 2>/dev/null
# LAST, because they have never run on a GPU and a protocol bug could wedge the device: the experimental paths
VITB_TEST_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_gemm_gpu.py tests/test_image_prep_gpu.py tests/test_model_gpu.py tests/test_kernels_gpu.py -q -k "register_layout or batch_loader or l16_geometry or key_split" > gpurun_out/${R}_pytest_exp.log 2>&1
echo "pytest experimental rc=$?"; tail -n 3 gpurun_out/${R}_pytest_exp.log
VITB_BENCH_EXPERIMENTAL=1 timeout 60 python tools/attn_bench.py > gpurun_out/${R}_attn_bench_exp.log 2>&1; echo "attn_bench experimental rc=$?"; tail -n 4 gpurun_out/${R}_attn_bench_exp.log

