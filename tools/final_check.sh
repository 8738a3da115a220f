#!/bin/bash
# Round-end validation on one B200: GPU tests, smoke, both bench arms, then ncu evidence (launch list of a step
# and full captures of the kernels that changed this round).  Every piece writes under gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/fc_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/fc_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fc_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/fc_smoke.log
timeout 300 python bench.py > gpurun_out/fc_bench_n1.json 2> gpurun_out/fc_bench_n1.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/fc_bench_n1.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fc_bench_ref.json 2> gpurun_out/fc_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/fc_bench_ref.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r01c.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/fc_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
for job in "attn_bwd:tools/attn_one.py::regex:attn_bwd" "gemm_fc1_gelu_dg:tools/gemm_one.py:fc1_gelu_dg:regex:vitb_gemm" \
           "gemm_dgrad_fc2_mul:tools/gemm_one.py:dgrad_fc2_mul:regex:vitb_gemm" "gemm_wgrad_pair:tools/gemm_one.py:wgrad_fc1_pair:regex:vitb_wgrad_pair"; do
  IFS=: read -r name script arg kflag kpat <<< "$job"
  timeout 120 ncu --set full --clock-control none --import-source on -k "$kflag:$kpat" -s 2 -c 1 -f -o "gpurun_out/${name}_r01c" \
    python "$script" $arg > "gpurun_out/fc_ncu_${name}.log" 2>&1; echo "ncu $name rc=$?"
done
ls -la gpurun_out/*_r01c* 2>/dev/null
