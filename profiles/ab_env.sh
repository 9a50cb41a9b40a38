#!/bin/bash
# profiles/ab_env.sh <lib> "ENV=.." "ENV=.." ...: the per-kernel table of one c2 frame under each environment setting
lib=$1; shift
for rep in 1 2; do
for e in "$@"; do
  echo "== $e (rep $rep)"
  env $e MOT_B200_LIB=$PWD/$lib timeout 120 python profiles/one_frame.py 6 2>&1 | egrep "k_rs_|k_cells|^\(" 
done; done
