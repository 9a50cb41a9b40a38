#!/bin/bash
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
TOPK=9 python profiles/exp_uf.py 16 4 -- "MOT_UF_MODE=2" "MOT_SORT_BIGTILE=0" "MOT_KEYS_HIST=0" "MOT_CSR_COMPACT=0" "MOT_SORT_BIGTILE=0 MOT_KEYS_HIST=0 MOT_CSR_COMPACT=0"
