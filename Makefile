# psac-b200 build: sm_100a only, in-tree shared library (travels to the GPU box with the snapshot)
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v --expt-relaxed-constexpr -I/usr/include
SRC := psac_b200/csrc
LIB := psac_b200/libpsacb200.so
HDRS := $(wildcard $(SRC)/*.cuh) include/psacb200.h

all: $(LIB) oracle

$(LIB): $(SRC)/engine.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRC)/engine.cu -lcudart -ldl 2> build_ptxas.log || (cat build_ptxas.log; exit 1)

oracle:
	$(MAKE) -s -f oracle/Makefile

clean:
	rm -f $(LIB) build_ptxas.log

.PHONY: all oracle clean
