#!/bin/bash
# One gpurun call that produces everything a round needs from the GPU box:
#   GPU tests, the bench line (ours + reference arm), the ncu launch list and one --set full capture per top kernel.
# usage (from the repo root, on the GPU box):  bash tools/gpu_round_check.sh [outdir]
OUT=${1:-gpurun_out/check}
mkdir -p $OUT
KREGEX='regex:pyramid|knn|mx_|gn_|match_filter|candidates|gather|unpack|gradient'
python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
# profiled command = the bench on the full workload, one timed step, small CPU sample
export VSB_CPU_SAMPLE_PAIRS=4
export VSB_BENCH_RAW_FRAMES=0     # the profiled command is the headline workload only (no from-raw-frames leg)
export VSB_BENCH_KNN_VARIANTS=0   # ... and no side-by-side matcher timing
python bench.py --steps 1 --warmup 1 > $OUT/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -k "$KREGEX" -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 > $OUT/ncu_l.log 2>&1
# steady-state launches of the three dominant kernels (skip the warm-up step's launches)
ncu --set full --clock-control none --import-source on -k 'regex:gn_solve_kernel|knn2_hamming_mx_bulk_kernel|pyramid16_kernel' \
    -s 3 -c 3 -o $OUT/prof_top -f python bench.py --steps 1 --warmup 1 > $OUT/ncu_f.log 2>&1
tail -3 $OUT/pytest.log
