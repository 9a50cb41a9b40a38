# Convenience targets; the same steps are what __graft_entry__.build() runs.
PKG   := multiple-object-tracking-lidar_b200
NVCC  ?= nvcc
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -shared

lib: $(PKG)/libmot_b200.so
$(PKG)/libmot_b200.so: $(wildcard $(PKG)/csrc/*.cu $(PKG)/csrc/*.cuh) include/mot_b200.h
	$(NVCC) $(NVCCFLAGS) -o $@ $(PKG)/csrc/mot_b200.cu

oracle:
	$(MAKE) -C oracle libmot_oracle.so

ref:            # the reference's own sources against stand-in headers; needs /root/reference
	$(MAKE) -C oracle _ref

test: lib oracle
	python -m pytest tests -q -m "not gpu"

test-gpu: lib oracle
	python -m pytest tests -q -m gpu

.PHONY: lib oracle ref test test-gpu
