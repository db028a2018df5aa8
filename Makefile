# Build of the B200-native TSDF hot path (sm_100a only) and of the CPU oracle.
#   make lib     -> tsdf_b200/libtsdf_b200.so   (CUDA kernels + C-ABI, the product)
#   make oracle  -> oracle/liboracle.so         (CPU restatement, test infrastructure)
#   make ref     -> oracle/_ref/libref_cuda.so  (reference's own .cu files, needs /root/reference)
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-Wall,-Wno-unused-function
CSRC      := tsdf_b200/csrc
OBJS      := $(CSRC)/integrate.o $(CSRC)/raycast.o $(CSRC)/misc.o $(CSRC)/volume.o $(CSRC)/mc.o $(CSRC)/bilateral.o $(CSRC)/exchange.o $(CSRC)/multi.o

all: lib oracle classes class_e2e

lib: tsdf_b200/libtsdf_b200.so

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/integrate_rigid.cuh $(CSRC)/mc_tables.h $(CSRC)/volume_internal.h include/tsdf_b200.h
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

# instrumented build of the raycast for tools/ray_iters.py (per-ray iteration classes, per-tile clocks); not the product
dbg: tsdf_b200/libtsdf_b200_dbg.so
tsdf_b200/libtsdf_b200_dbg.so: $(OBJS)
	$(NVCC) $(NVCCFLAGS) -DTSDF_RAY_DEBUG -c $(CSRC)/raycast.cu -o $(CSRC)/raycast_dbg.o
	$(NVCC) $(ARCH) -shared -o $@ $(subst raycast.o,raycast_dbg.o,$(OBJS))

# tuning variants of the raycast (A/B runs through TSDF_B200_LIB): make variant NAME=minb5 RAYFLAGS=-DTSDF_RAY_MINB=5
variant: $(OBJS)
	$(NVCC) $(NVCCFLAGS) $(RAYFLAGS) -c $(CSRC)/raycast.cu -o $(CSRC)/raycast_$(NAME).o
	$(NVCC) $(ARCH) -shared -o tsdf_b200/libtsdf_b200_$(NAME).so $(subst raycast.o,raycast_$(NAME).o,$(OBJS))

tsdf_b200/libtsdf_b200.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

# Drop-in C++ class layer (TSDFVolume / Camera / GPURaycaster / loaders / PNG / PLY / marching cubes) over the C-ABI:
# host-only code, no CUDA; links against libtsdf_b200.so.  Eigen is replaced by tsdf_b200/compat unless EIGEN_INC is given.
HOST      := tsdf_b200/host
HOSTSRC   := $(wildcard $(HOST)/*.cpp)
HOSTOBJ   := $(HOSTSRC:.cpp=.o)
EIGEN_INC ?= tsdf_b200/compat
# TSDF_B200_PINNED_EIGEN: Dynamic matrices of the Eigen stand-in take their storage from tsdf_b200_host_alloc (pinned pool)
CXXFLAGS  := -O2 -std=c++14 -fPIC -Wall -Wno-unused-function -DTSDF_B200_PINNED_EIGEN -I$(EIGEN_INC) -I/usr/local/cuda/include

classes: tsdf_b200/libtsdf_b200_classes.so

$(HOST)/%.o: $(HOST)/%.cpp $(wildcard tsdf_b200/include/*.hpp) $(wildcard tsdf_b200/compat/*.hpp) include/tsdf_b200.h
	g++ $(CXXFLAGS) -c $< -o $@

tsdf_b200/libtsdf_b200_classes.so: $(HOSTOBJ) tsdf_b200/libtsdf_b200.so
	g++ -shared -o $@ $(HOSTOBJ) -Ltsdf_b200 -ltsdf_b200 -lz -Wl,-rpath,'$$ORIGIN'

# The reference's own driver, compiled UNCHANGED against this repo's headers: build/dropin/Tools/kinfu.cpp is a symlink to
# the reference file, build/dropin/include a symlink to tsdf_b200/include, so its "../include/..." includes resolve here.
REF_SRC ?= /root/reference/src
kinfu: classes
	mkdir -p build/dropin/Tools
	ln -sfn $(abspath tsdf_b200/include) build/dropin/include
	ln -sf $(REF_SRC)/Tools/kinfu.cpp build/dropin/Tools/kinfu.cpp
	g++ $(CXXFLAGS) -include cstring -include cstdlib -include cstdint -o build/kinfu build/dropin/Tools/kinfu.cpp \
	    -Ltsdf_b200 -ltsdf_b200_classes -ltsdf_b200 -Wl,-rpath,$(abspath tsdf_b200)

# End-to-end timing of the class layer the way kinfu.cpp drives it (tools/class_e2e.cpp; bench.py runs it when it exists)
class_e2e: classes
	mkdir -p build
	g++ $(CXXFLAGS) -o build/class_e2e tools/class_e2e.cpp -Ltsdf_b200 -ltsdf_b200_classes -ltsdf_b200 -Wl,-rpath,$(abspath tsdf_b200)

oracle: oracle/liboracle.so

oracle/liboracle.so: oracle/tsdf_oracle.c tsdf_b200/csrc/mc_tables.h
	gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared -Wall -o $@ $< -lm

ref:
	bash oracle/build_ref.sh

clean:
	rm -f $(CSRC)/*.o $(HOST)/*.o tsdf_b200/libtsdf_b200.so tsdf_b200/libtsdf_b200_classes.so oracle/liboracle.so

.PHONY: all lib oracle ref clean classes kinfu variant dbg class_e2e
