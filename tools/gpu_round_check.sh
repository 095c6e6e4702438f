#!/bin/bash
# One gpurun call that produces everything a round needs from the GPU box:
#   GPU tests, the bench line (ours + reference arm), the ncu launch list and one --set full capture per top kernel.
# usage (from the repo root, on the GPU box):  bash tools/gpu_round_check.sh [outdir]
OUT=${1:-gpurun_out/check}
mkdir -p $OUT
KREGEX='regex:pyramid|knn|mx_|gn_|match_filter|candidates|gather|unpack|gradient|l2_'
python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
# profiled command = the headline workload (configs[1]) through the tracker: one warm-up pass + timed passes
python tools/leg_once.py 1 3 > $OUT/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -k "$KREGEX" -s 7 -c 21 --csv --log-file $OUT/launches.csv \
    python tools/leg_once.py 1 3 > $OUT/ncu_l.log 2>&1
python tools/launch_table.py $OUT/launches.csv > $OUT/launches.txt
# steady-state launches of every kernel of the step (skip the warm-up pass: 7 launches)
ncu --set full --clock-control none --import-source on -k 'regex:gn_track|mx_bulk|candidates|pyramid16|mx_expand|match_filter' \
    -s 7 -c 7 -o $OUT/prof_step -f python tools/leg_once.py 1 1 > $OUT/ncu_f.log 2>&1
tail -3 $OUT/pytest.log
# the ORB front end (from raw frames): per-kernel table over two 500-frame calls, and one --set full capture of its main kernels
python tools/orb_once.py > $OUT/orb_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k 'regex:orb_|fast_' --csv --log-file $OUT/orb_metrics.csv python tools/orb_once.py > $OUT/ncu_orb.log 2>&1
python tools/orb_kernel_table.py $OUT/orb_metrics.csv > $OUT/orb_kernels.txt
ncu --set full --clock-control none --import-source on -k 'regex:fast_score_kernel|orb_blur_kernel|orb_resize_tile_kernel|fast_count16_kernel|orb_describe_kernel' \
    -s 5 -c 5 -o $OUT/orb_full -f python tools/orb_once.py > $OUT/ncu_orb_f.log 2>&1
