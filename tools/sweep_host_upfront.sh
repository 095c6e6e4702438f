#!/bin/bash
# e2e (host-buffer) pass of the bench with the side inputs (descriptors, key points, counts, priors) uploaded per chunk
# (VSB_HOST_UPFRONT=0) or once per sequence (default).   -> gpurun_out/hu_<0|1>.json
export VSB_CPU_SAMPLE_PAIRS=16 VSB_BENCH_RAW_FRAMES=0 VSB_BENCH_KNN_VARIANTS=0
mkdir -p gpurun_out
for s in 0 1 0 1; do
  VSB_HOST_UPFRONT=$s python bench.py --steps 5 --warmup 3 > gpurun_out/hu_$s.json 2> gpurun_out/hu_$s.err
  python -c "
import json; d=json.loads(open('gpurun_out/hu_$s.json').read().strip().splitlines()[-1]); e=d['e2e']; print($s, round(e['ms_per_step'],3), round(e['value']), e['h2d_bytes_per_step'], round(e['h2d_gbs_achieved'],2), round(e['h2d_gbs_plain_copy'],2), e['matches_device_path'], round(d['ms_per_step'],3))"
done
