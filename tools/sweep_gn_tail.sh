#!/bin/bash
# GN tail-launch sweep: pairs of the tail launch x threads per pair (BASELINE configs[1], 1999 pairs)
for tp in 0 100 223 400 600; do for tt in 256 512 1024; do
  echo "tail_pairs=$tp tail_threads=$tt: $(VSB_GN_TAIL_PAIRS=$tp VSB_GN_TAIL_THREADS=$tt python tools/kbench_gn.py 2000 1:0:1:8192 2>&1 | tail -1)"
  [ $tp = 0 ] && break
done; done
