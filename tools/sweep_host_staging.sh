#!/bin/bash
# e2e (host-buffer) pass of the bench with the frames uploaded straight into the pyramid (strided copy, default) or through a
# contiguous staging block (VSB_HOST_STAGING=1).   -> gpurun_out/hs_<0|1>.json
export VSB_CPU_SAMPLE_PAIRS=16 VSB_BENCH_RAW_FRAMES=0 VSB_BENCH_KNN_VARIANTS=0
mkdir -p gpurun_out
for s in 0 1 0 1; do
  VSB_HOST_STAGING=$s python bench.py --steps 5 --warmup 3 > gpurun_out/hs_$s.json 2> gpurun_out/hs_$s.err
  python -c "
import json; d=json.loads(open('gpurun_out/hs_$s.json').read().strip().splitlines()[-1]); e=d['e2e']; print($s, round(e['ms_per_step'],3), round(e['value']), round(e['h2d_gbs_achieved'],2), round(e['h2d_gbs_plain_copy'],2), e['matches_device_path'], round(d['ms_per_step'],3))"
done
