#!/bin/bash
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
MOT_UF_MODE=2 TOPK=5 python profiles/kernels_of.py c2frame
MOT_UF_MODE=2 MOT_CELL_DENSE=0 TOPK=3 python profiles/kernels_of.py c2frame
TOPK=4 python profiles/exp_uf.py 16 4 -- "MOT_UF_MODE=2" "MOT_CELL_DENSE=0"
MOT_UF_MODE=2 TOPK=5 python profiles/kernels_of.py c4:1.0
