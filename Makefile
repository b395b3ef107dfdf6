# Builds dgraph_b200/libdgx.so (sm_100a only) and the CPU oracle.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS ?= -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-maybe-uninitialized -Xptxas -v
CSRC := dgraph_b200/csrc
HDRS := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.hpp) include/dgx.h

all: dgraph_b200/libdgx.so oracle/liboracle.so

dgraph_b200/libdgx.so: $(CSRC)/dgx_api.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/dgx_api.cu -lcudart

oracle/liboracle.so: oracle/oracle.c oracle/oracle.h
	$(MAKE) -C oracle liboracle.so

clean:
	rm -f dgraph_b200/libdgx.so oracle/liboracle.so
