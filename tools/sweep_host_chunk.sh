#!/bin/bash
# e2e (host-buffer) pass of the bench for several H2D/compute pipeline chunk sizes (pairs per chunk).
# usage: bash tools/sweep_host_chunk.sh "250 125 64"   -> gpurun_out/hc_<chunk>.json
export VSB_CPU_SAMPLE_PAIRS=16 VSB_BENCH_RAW_FRAMES=0 VSB_BENCH_KNN_VARIANTS=0
mkdir -p gpurun_out
for c in ${1:-250 125 64}; do
  VSB_BENCH_HOST_CHUNK=$c python bench.py --steps 5 --warmup 3 > gpurun_out/hc_$c.json 2> gpurun_out/hc_$c.err
  python -c "
import json; d=json.loads(open('gpurun_out/hc_$c.json').read().strip().splitlines()[-1]); e=d['e2e']; print($c, round(e['ms_per_step'],3), round(e['value']), round(e['h2d_gbs_achieved'],2), round(e['h2d_gbs_plain_copy'],2), e['matches_device_path'], round(d['ms_per_step'],3))"
done
