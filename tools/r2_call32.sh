#!/bin/bash
# shapes the key-block attention backward was written for: ViT-H/14 training, 384 px training; c4 inference re-measured
TAG=${ROUND_TAG:-r02r}
mkdir -p gpurun_out
for cfg in h14train b16_384; do
  for long in 1 0; do
    VITB_ATTN_BWD_LONG=$long timeout 400 python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/${TAG}_${cfg}_long$long.json 2> gpurun_out/${TAG}_${cfg}_long$long.err
    echo "$cfg long=$long rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_${cfg}_long$long.json'));print('%.1f img/s %.2f ms/step step %.3f of sustained' % (d['value'], d['ms_per_step'], d['roofline_step']['frac_of_sustained']))")"
  done
done
timeout 400 python bench.py --config c4 --steps 10 --warmup 3 > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err
echo "c4 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_c4.json'));print('%.1f img/s %.2f ms' % (d['value'], d['ms_per_step']))")"
timeout 400 python bench.py --config c3 --steps 10 --warmup 3 > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
echo "c3 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_c3.json'));print('%.1f img/s %.2f ms' % (d['value'], d['ms_per_step']))")"
timeout 400 python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/${TAG}_c5.json 2> gpurun_out/${TAG}_c5.err
echo "c5 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_c5.json'));print('%.1f img/s %.2f ms' % (d['value'], d['ms_per_step']))")"
