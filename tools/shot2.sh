#!/bin/bash
mkdir -p gpurun_out
timeout 40 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/s4_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"
tail -3 gpurun_out/s4_pytest.log
timeout 30 python bench.py --steps 8 --warmup 3 > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/s4_bench.json
