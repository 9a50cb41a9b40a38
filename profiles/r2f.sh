#!/bin/bash
PK=multiple-object-tracking-lidar_b200
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2f_pytest.log
TOPK=9 python profiles/exp_uf.py 16 4 -- "MOT_UF_MODE=1" "MOT_UF_MODE=2" "MOT_UF_PHASES=1,30" "MOT_UF_PHASES=7,24" "MOT_UF_PHASES=1,2,28" "MOT_UF_WBLOCKS=4" "MOT_UF_WBLOCKS=12" "MOT_UF_XMODE=0" > gpurun_out/r2f_exp.log 2>&1
MOT_B200_LIB=$PWD/$PK/libmot_b200_stats1.so python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2f_stats1.log 2>&1
python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2f_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_uf_survivors|k_uf_walk" -s 18 -c 6 -o gpurun_out/r2f_uf python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2f_ncu.log 2>&1
cat gpurun_out/r2f_pytest.log gpurun_out/r2f_exp.log; head -3 gpurun_out/r2f_stats1.log
tail -3 gpurun_out/r2f_ncu.log
