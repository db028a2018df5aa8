# Build of the B200-native TSDF hot path (sm_100a only) and of the CPU oracle.
#   make lib     -> tsdf_b200/libtsdf_b200.so   (CUDA kernels + C-ABI, the product)
#   make oracle  -> oracle/liboracle.so         (CPU restatement, test infrastructure)
#   make ref     -> oracle/_ref/libref_cuda.so  (reference's own .cu files, needs /root/reference)
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-Wall,-Wno-unused-function
CSRC      := tsdf_b200/csrc
OBJS      := $(CSRC)/integrate.o $(CSRC)/raycast.o $(CSRC)/misc.o $(CSRC)/volume.o

all: lib oracle

lib: tsdf_b200/libtsdf_b200.so

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/integrate_rigid.cuh include/tsdf_b200.h
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

tsdf_b200/libtsdf_b200.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

oracle: oracle/liboracle.so

oracle/liboracle.so: oracle/tsdf_oracle.c
	gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared -Wall -o $@ $< -lm

ref:
	bash oracle/build_ref.sh

clean:
	rm -f $(CSRC)/*.o tsdf_b200/libtsdf_b200.so oracle/liboracle.so

.PHONY: all lib oracle ref clean
