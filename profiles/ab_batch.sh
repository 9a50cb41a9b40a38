#!/bin/bash
# A/B of library builds on one box, batch mode and single-frame mode: profiles/ab_batch.sh ab/base.so ab/new.so ...
for rep in 1 2; do
for lib in "$@"; do
  echo "== $lib (rep $rep)"
  MOT_B200_LIB=$PWD/$lib timeout 200 python profiles/batch_kernels.py 16 6 2>&1 | egrep "k_uf_sparse|k_cells_write|kernel time"
  MOT_B200_LIB=$PWD/$lib timeout 120 python profiles/one_frame.py 6 2>&1 | egrep "k_uf_sparse|k_cells_write"
done; done
