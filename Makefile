# Builds the C-ABI library (CUDA, sm_100a), the plain-C oracle restatement and, when /root/reference
# is present, the reference itself (oracle/_ref).  Used by __graft_entry__.build().
NVCC      ?= nvcc
CXX       ?= g++
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC
PKG       := vc2_reference_b200
CSRC      := $(PKG)/csrc
LIB       := $(PKG)/libvc2b200.so
OBJS      := $(CSRC)/dwt_fwd.o $(CSRC)/dwt_inv.o $(CSRC)/slices.o $(CSRC)/cabi.o

ORACLE    := oracle/_build/libvc2oracle.so

all: $(LIB) $(ORACLE)

$(ORACLE): oracle/vc2_oracle.c
	mkdir -p oracle/_build
	$(CC) -O2 -std=c99 -fPIC -shared $< -o $@ -lm

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/*.cuh include/vc2_cabi.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

# the lifting kernels are one source compiled twice (forward / inverse) so that the two halves build in parallel
$(CSRC)/dwt_fwd.o: $(CSRC)/dwt.cu $(CSRC)/*.cuh include/vc2_cabi.h
	$(NVCC) $(NVFLAGS) -DVC2_DWT_PART=1 -c $< -o $@
$(CSRC)/dwt_inv.o: $(CSRC)/dwt.cu $(CSRC)/*.cuh include/vc2_cabi.h
	$(NVCC) $(NVFLAGS) -DVC2_DWT_PART=2 -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

clean:
	rm -f $(OBJS) $(LIB)
.PHONY: all clean
