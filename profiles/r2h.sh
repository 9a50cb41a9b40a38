#!/bin/bash
PK=multiple-object-tracking-lidar_b200
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2h_pytest.log
for mb in 6 10 12; do TOPK=3 MOT_B200_LIB=$PWD/$PK/libmot_b200_f$mb.so python profiles/exp_uf.py 16 4 -- "MOT_UF_FBLOCKS=$mb" "MOT_UF_FBLOCKS=$mb MOT_UF_LIGHT=16" > gpurun_out/r2h_f$mb.log 2>&1; done
cat gpurun_out/r2h_pytest.log; grep -A4 "===" gpurun_out/r2h_f*.log
