#!/bin/bash
# A/B of two builds of the library on one box: profiles/ab.sh ab/base.so ab/e1.so ...  (k_uf_sparse + total per frame)
for rep in 1 2; do
for lib in "$@"; do
  echo "== $lib (rep $rep)"
  MOT_B200_LIB=$PWD/$lib timeout 120 python profiles/one_frame.py 6 2>&1 | egrep "k_uf_sparse|k_uf_dense|^\(" 
done; done
