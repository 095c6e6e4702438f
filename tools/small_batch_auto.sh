for n in 2 10 19 28 38 75 100 126; do python tools/leg_once.py 1 10 $n 2>&1 | tail -1 | sed 's/kernels_ms.*gn_solve/gn_solve/'; done
