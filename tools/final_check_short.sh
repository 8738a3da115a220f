#!/bin/bash
# Short round-end validation on one B200 (no ncu): GPU tests, smoke, both bench arms, stock-torch comparator.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/fs_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/fs_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fs_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/fs_smoke.log
timeout 200 python bench.py > gpurun_out/fs_bench_n1.json 2> gpurun_out/fs_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/fs_bench_n1.json
timeout 100 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fs_bench_ref.json 2> gpurun_out/fs_bench_ref.err; echo "ref rc=$?"; cut -c1-120 gpurun_out/fs_bench_ref.json
timeout 100 python bench.py --impl reference --ref-device cuda --ref-autocast --steps 5 --warmup 2 > gpurun_out/fs_stock_torch.json 2> gpurun_out/fs_stock_torch.err; echo "stock rc=$?"; cut -c1-200 gpurun_out/fs_stock_torch.json
