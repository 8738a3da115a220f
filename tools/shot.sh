#!/bin/bash
# One short GPU call: the GPU test suite, the epilogue / bandwidth-kernel A/B, one bench line.  Each leg writes its own log.
mkdir -p gpurun_out
timeout 60 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/s3_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"
tail -4 gpurun_out/s3_pytest.log
timeout 30 python tools/epi_ab.py > gpurun_out/s3_ab.log 2>&1; echo "ab rc=$?"
tail -12 gpurun_out/s3_ab.log
timeout 40 python bench.py --steps 10 --warmup 3 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?"
cut -c1-330 gpurun_out/s3_bench.json
