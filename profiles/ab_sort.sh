# A/B of library builds of the radix scatter (round 2): the variants were built with
#   nvcc ... -DRS_BIG_THREADS_V=256|512|1024 [-DRS_SCATTER_CTAS_BIG=n] [-DRS_RANK_LDST] -o ab/<name>.so csrc/mot_b200.cu
# and compared on one box; results: gpurun_out/s3_ab_sort*.log of the session, summarised in DESIGN.md section 7.
for rep in 1 2; do for lib in t256 t512 t1024; do echo "== $lib (rep $rep)"; MOT_B200_LIB=$PWD/ab/$lib.so TOPK=8 timeout 200 python profiles/batch_kernels.py 16 5 2>&1 | egrep "kernel time|k_rs_scatter|k_rs_hist"; done; done
