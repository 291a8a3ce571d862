# Builds the product library (CUDA, sm_100a only) and the oracle's C restatement (CPU checker).
NVCC      ?= /usr/local/cuda/bin/nvcc
CUDA_LIB  ?= /usr/local/cuda/lib64
PKG       := homogenization.jl_b200
CSRC      := $(PKG)/csrc
LIB       ?= $(PKG)/libhmg_b200.so
BUILD     ?= build
NVFLAGS   := $(HMG_EXTRA) -DHMG_NVTX=1 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
             -Xcompiler -fPIC,-Wall,-Wno-unused-function -Iinclude -diag-suppress 20014
SRCS      := $(CSRC)/api.cu $(CSRC)/kernels.cu $(CSRC)/field.cu $(CSRC)/reference.cpp $(CSRC)/topology.cpp $(CSRC)/introspect.cpp
HDRS      := include/hmg.h $(CSRC)/hmg_host.hpp $(CSRC)/kernels.cuh $(CSRC)/lattice.hpp $(CSRC)/apply_core.cuh
OBJS      := $(patsubst $(CSRC)/%,$(BUILD)/%.o,$(SRCS))

all: $(LIB) oracle checked

$(LIB): $(OBJS)
	@mkdir -p $(dir $@)
	$(NVCC) -shared -o $@ $(OBJS) -L$(CUDA_LIB) -lcusolver -lcublas -ldl \
	    -Xlinker -rpath=$(CUDA_LIB)

$(BUILD)/%.cu.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; false)

$(BUILD)/%.cpp.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

# the library with device-side bounds assertions (this pool has no compute-sanitizer): variants/libhmg_checked.so
checked:
	$(MAKE) BUILD=build/checked LIB=variants/libhmg_checked.so HMG_EXTRA="-DHMG_BOUNDS" variants/libhmg_checked.so

oracle:
	@if [ -f oracle/c/Makefile ]; then $(MAKE) -C oracle/c; fi

clean:
	rm -rf build $(LIB)

.PHONY: all oracle clean checked
