#!/bin/bash
PK=multiple-object-tracking-lidar_b200
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2d_pytest.log
for mb in 3 4 5 6; do
  TOPK=4 MOT_B200_LIB=$PWD/$PK/libmot_b200_mb$mb.so python profiles/exp_uf.py 16 4 -- "MOT_UF_MODE=1" "MOT_UF_MODE=2 MOT_UF_XBLOCKS=$mb" > gpurun_out/r2d_mb$mb.log 2>&1
done
python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2d_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_uf_cross|k_cell_local" -s 6 -c 2 -o gpurun_out/r2d_uf python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2d_ncu.log 2>&1
cat gpurun_out/r2d_pytest.log; for mb in 3 4 5 6; do echo "## mb$mb"; grep -A6 "MODE=2" gpurun_out/r2d_mb$mb.log; done
tail -3 gpurun_out/r2d_ncu.log
