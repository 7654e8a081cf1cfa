# Builds libvct_cuda.so (the C-ABI drop-in boundary, sm_100a only) in-tree, the C++ host layer
# and the CPU oracle.  `make` = everything; __graft_entry__.build() calls this.
NVCC     := /usr/local/cuda/bin/nvcc
CXX      := /usr/bin/g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
CSRC     := voxel_cone_tracing_b200/csrc
HOST     := voxel_cone_tracing_b200/host
OUT      := voxel_cone_tracing_b200
NVFLAGS  := -O3 -std=c++17 $(ARCH) -lineinfo -Iinclude -I$(CSRC) -Xcompiler -fPIC,-Wall,-ffp-contract=off -ccbin $(CXX)
# bit-exact units (same arithmetic as the oracle: no FMA contraction)
EXACT    := -fmad=false
OBJS     := $(CSRC)/capi.o $(CSRC)/tex3d.o $(CSRC)/voxelize.o $(CSRC)/mipmap.o $(CSRC)/gbuffer.o $(CSRC)/cone_trace.o $(CSRC)/peer.o
HDRS     := include/vct/vct_c.h $(CSRC)/vct_internal.cuh $(CSRC)/raster.cuh $(CSRC)/mip_arith.cuh

all: $(OUT)/libvct_cuda.so oracle host hosttest

$(CSRC)/voxelize.o: $(CSRC)/voxelize.cu $(HDRS)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(CSRC)/mipmap.o: $(CSRC)/mipmap.cu $(HDRS)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(CSRC)/gbuffer.o: $(CSRC)/gbuffer.cu $(HDRS)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(CSRC)/cone_trace.o: $(CSRC)/cone_trace.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@
$(CSRC)/capi.o: $(CSRC)/capi.cu $(HDRS)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(CSRC)/tex3d.o: $(CSRC)/tex3d.cu $(HDRS)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(CSRC)/peer.o: $(CSRC)/peer.cu $(HDRS)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@

$(OUT)/libvct_cuda.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart shared

oracle:
	$(MAKE) -C oracle all

host: $(OUT)/libvct_cuda.so
	@if [ -f $(HOST)/Makefile ]; then $(MAKE) -C $(HOST); fi

# host (g++) build of the arithmetic headers the kernels share, for the CPU test suite (tests/test_mip_arith.py)
hosttest: tests/native/libvct_hosttest.so
tests/native/libvct_hosttest.so: tests/native/mip_arith_host.cpp $(CSRC)/mip_arith.cuh
	$(CXX) -O2 -std=c++17 -ffp-contract=off -fPIC -shared -I$(CSRC) -o $@ $<

clean:
	rm -f $(OBJS) $(OUT)/libvct_cuda.so
	$(MAKE) -C oracle clean
.PHONY: all oracle host hosttest clean
