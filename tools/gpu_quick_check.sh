#!/bin/bash
# Quick GPU check after a kernel change: the GPU tests, a lean bench line, one --set full capture of one kernel.
# usage: KERNEL=candidates_kernel bash tools/gpu_quick_check.sh
OUT=${OUT:-gpurun_out/quick}
mkdir -p ${OUT}
python -m pytest tests -m gpu -x -q > ${OUT}/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 ${OUT}/pytest.log
export VSB_CPU_SAMPLE_PAIRS=64 VSB_BENCH_RAW_FRAMES=0 VSB_BENCH_KNN_VARIANTS=0 VSB_BENCH_CONFIGS=none
python bench.py --steps 5 --warmup 3 > ${OUT}/bench.json 2> ${OUT}/bench.err; echo "bench rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:${KERNEL:-candidates_kernel}" -s 1 -c 1 -o ${OUT}/prof_cand -f python tools/leg_once.py 1 1 > ${OUT}/ncu.log 2>&1
python -c "
import json; d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], [(k['kernel'], round(k['ms_per_step'],4)) for k in d['kernels']], d['cpu_baseline']['max_abs_pose_diff_vs_gpu'])"
