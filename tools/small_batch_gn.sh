#!/bin/bash
# Small batches of configs[1] through the tracker: one block per pair (cluster=0) against one pair per thread-block cluster
# of 2 / 4 / 8 blocks of 256 / 512 threads.  Usage (GPU box): bash tools/small_batch_gn.sh [frames ...]
for n in ${@:-2 10 19 38 75 126 251}; do
  for mode in "0 512" "2 256" "2 512" "4 256" "4 512" "8 256" "8 512"; do
    set -- $mode
    echo -n "cluster=$1 cthreads=$2  "
    VSB_GN_CLUSTER=$1 VSB_GN_CLUSTER_THREADS=$2 python tools/leg_once.py 1 10 $n 2>&1 | tail -1 | sed 's/kernels_ms.*gn_solve/gn_solve/'
  done
done
