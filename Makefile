# Builds libdysb200.so (the C-ABI CUDA library) in-tree for sm_100a.
PKG      := recognizing-speech-dysfluencies-in-stuttering_b200
CSRC     := $(PKG)/csrc
NVCC     ?= /usr/local/cuda/bin/nvcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr $(VARIANT_FLAGS)
SRCS     := $(CSRC)/dys_api.cu $(CSRC)/dys_tables.cu $(CSRC)/dys_features.cu $(CSRC)/dys_denoise.cu $(CSRC)/dys_cmvn.cu $(CSRC)/dys_qc.cu $(CSRC)/dys_profile.cu $(CSRC)/dys_resample.cu
OBJS     := $(SRCS:.cu=.o)
LIB      ?= $(PKG)/libdysb200.so

all: $(LIB)

$(CSRC)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/dysfluency_b200.h
	$(NVCC) $(NVFLAGS) $(PTXAS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -lcudart

clean:
	rm -f $(OBJS) $(LIB)

.PHONY: all clean
