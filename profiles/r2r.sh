#!/bin/bash
TOPK=3 python profiles/exp_uf.py 16 4 -- "MOT_UF_MODE=2" "MOT_UF_ROWINNER=1"
MOT_UF_MODE=2 MOT_UF_ROWINNER=1 TOPK=3 python profiles/kernels_of.py c2frame
EXP_WORKLOAD=c3 TOPK=3 python profiles/exp_uf.py 64 3 -- "MOT_UF_MODE=2" "MOT_UF_ROWINNER=1"
