#!/bin/bash
# A/B of library builds on one box in batch mode: profiles/ab_batch.sh ab/base.so ab/new.so ...   (PAT = egrep pattern)
PAT=${PAT:-"k_uf_sparse|kernel time"}
for rep in 1 2; do
for lib in "$@"; do
  echo "== $lib (rep $rep)"
  MOT_B200_LIB=$PWD/$lib TOPK=40 timeout 200 python profiles/batch_kernels.py 16 6 2>&1 | egrep "$PAT"
done; done
