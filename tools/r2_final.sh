#!/bin/bash
# Round-end validation on one B200: GPU tests, smoke, both bench arms, the launch list of a step (ncu, DRAM bytes), and
# ncu captures (limited sections, small reports) of the kernels written late in the round.
cd "$(dirname "$0")/.." || exit 1
TAG=${ROUND_TAG:-r02z}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/${TAG}_bench_n1.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-260 gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
SECTIONS="--section SpeedOfLight --section WarpStateStats --section SchedulerStats --section LaunchStats --section MemoryWorkloadAnalysis --section Occupancy"
ATTN_SHAPE=32,577,12,64 timeout 200 ncu $SECTIONS --clock-control none -k regex:attn_bwd_tc -s 1 -c 1 -f -o gpurun_out/${TAG}_attn_bwd_long \
  python tools/attn_one.py > gpurun_out/${TAG}_ncu_attn_long.log 2>&1; echo "ncu attn long rc=$?"
ATTN_SHAPE=32,257,16,80 timeout 200 ncu $SECTIONS --clock-control none -k regex:attn_bwd_tc_wide -s 1 -c 1 -f -o gpurun_out/${TAG}_attn_bwd_wide \
  python tools/attn_one.py > gpurun_out/${TAG}_ncu_attn_wide.log 2>&1; echo "ncu attn wide rc=$?"
timeout 200 ncu $SECTIONS --clock-control none -k regex:attn_bwd_ws -s 1 -c 1 -f -o gpurun_out/${TAG}_attn_bwd_ws \
  python tools/attn_one.py > gpurun_out/${TAG}_ncu_attn_ws.log 2>&1; echo "ncu attn ws rc=$?"
for r in attn_bwd_long attn_bwd_wide attn_bwd_ws; do
  ncu -i gpurun_out/${TAG}_$r.ncu-rep --page details --csv > gpurun_out/${TAG}_$r.details.csv 2>/dev/null
done
ls -la gpurun_out/${TAG}_*.ncu-rep 2>/dev/null
