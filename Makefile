# Build of the B200-native CFRK hot path.
#   make            -> cfrk_b200/lib/libcfrk_b200.so (C ABI, include/cfrk_b200.h) + bin/cfrk (CLI)
#   make oracle     -> oracle/_build/liboracle.so (+ oracle/_ref/* when /root/reference exists)
NVCC    ?= nvcc
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unknown-pragmas -Xptxas -v
CSRC    := cfrk_b200/csrc
OBJ     := build/obj
LIB     := cfrk_b200/lib/libcfrk_b200.so
CLI     := bin/cfrk

CU_SRCS  := $(CSRC)/kernels.cu $(CSRC)/dense_lane.cu $(CSRC)/row_pairs.cu $(CSRC)/sparse.cu $(CSRC)/hist_split.cu $(CSRC)/hist_reduce.cu $(CSRC)/fasta_scan.cu $(CSRC)/api.cu $(CSRC)/compat_shim.cu $(CSRC)/runfile.cu
CU_OBJS  := $(patsubst $(CSRC)/%.cu,$(OBJ)/%.o,$(CU_SRCS))
HDRS     := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/cfrk_b200.h

all: $(LIB) $(CLI)

$(OBJ)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || { cat $(OBJ)/$*.ptxas.log; exit 1; }

$(LIB): $(CU_OBJS)
	@mkdir -p cfrk_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(CU_OBJS) -lpthread -lz

$(CLI): $(CSRC)/cli_main.cpp $(LIB)
	@mkdir -p bin
	g++ -O2 -std=c++17 -Wall -Iinclude -o $@ $(CSRC)/cli_main.cpp -Lcfrk_b200/lib -lcfrk_b200 -Wl,-rpath,'$$ORIGIN/../cfrk_b200/lib'

oracle:
	$(MAKE) -C oracle all
	@if [ -d /root/reference/src ]; then $(MAKE) -C oracle ref; fi

clean:
	rm -rf build cfrk_b200/lib bin

.PHONY: all oracle clean
