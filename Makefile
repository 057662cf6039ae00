# Builds libmacb200.so (sm_100a only) in-tree.  `python -c "import __graft_entry__ as g; g.build()"` runs this.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr
SRC := mac_b200/csrc/api.cu mac_b200/csrc/tridiag.cpp
HDR := mac_b200/csrc/kernels.cuh mac_b200/csrc/tridiag.h include/macb200.h
OUT := mac_b200/libmacb200.so

all: $(OUT)

$(OUT): $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRC) -ldl

ptxas-info:
	$(NVCC) $(NVFLAGS) -Xptxas -v -shared -o /tmp/macb_ptxas.so $(SRC) 2>&1 | grep -E 'Compiling|registers|spill' 

clean:
	rm -f $(OUT)
.PHONY: all clean ptxas-info

# debug variant with per-phase clock64 stamps inside the persistent Lanczos kernel (tools/ptiming.py)
debugvec: $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -DMACB_DEBUG_VEC -shared -o mac_b200/libmacb200_debug.so $(SRC)

timing: $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -DMACB_PTIMING -shared -o mac_b200/libmacb200_timing.so $(SRC)
