set -x
mkdir -p gpurun_out/s4
python -m pytest tests -m gpu -x -q > gpurun_out/s4/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4/pytest.log
python bench.py > gpurun_out/s4/bench.json 2> gpurun_out/s4/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s4/bench_ref.json 2> gpurun_out/s4/bench_ref.err
export VSB_BENCH_FRAMES=301 VSB_CPU_SAMPLE_PAIRS=4
python bench.py --steps 2 --warmup 1 > gpurun_out/s4/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s4/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/s4/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gn_solve_kernel|knn2_hamming_tc_kernel|pyramid_kernel" -s 6 -c 3 -o gpurun_out/s4/prof_top -f python bench.py --steps 2 --warmup 1 > gpurun_out/s4/ncu_f.log 2>&1
tail -3 gpurun_out/s4/pytest.log
