# Builds the C-ABI library (CUDA, sm_100a), the plain-C oracle restatement and, when /root/reference
# is present, the reference itself (oracle/_ref).  Used by __graft_entry__.build().
NVCC      ?= nvcc
CXX       ?= g++
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 -lineinfo -diag-suppress 177 $(ARCH) -Xcompiler -fPIC $(EXTRA)
PKG       := vc2_reference_b200
CSRC      := $(PKG)/csrc
LIB       := $(PKG)/libvc2b200.so
OBJS      := $(CSRC)/dwt_fwd.o $(CSRC)/dwt_inv.o $(CSRC)/dwt_tile_fwd.o $(CSRC)/dwt_tile_inv.o $(CSRC)/slices.o $(CSRC)/cabi.o

ORACLE    := oracle/_build/libvc2oracle.so

HOSTLIB   := $(PKG)/libvc2host.so
BIN       := $(PKG)/bin
CXXFLAGS  := -O2 -std=c++14 -fPIC -Wall -Wno-comment -Iinclude -Ihost

all: $(LIB) $(ORACLE) $(HOSTLIB) $(BIN)/EncodeStream $(BIN)/DecodeStream $(BIN)/DecodeFrame $(BIN)/test_library_mirror $(BIN)/test_pipeline_host $(BIN)/test_vlc_host

# C++ host layer: the Library mirror (include/vc2/*.h) and the drop-in command lines, over the C-ABI
$(HOSTLIB): host/vc2_library.cpp host/vc2_stream.cpp include/vc2/*.h include/vc2_cabi.h include/vc2_host.h $(LIB)
	$(CXX) $(CXXFLAGS) -shared host/vc2_library.cpp host/vc2_stream.cpp -o $@ -L$(PKG) -lvc2b200 -Wl,-rpath,'$$ORIGIN'
$(BIN)/%: host/%.cpp host/cmdline.h host/pipeline.h include/vc2/*.h include/vc2_cabi.h $(HOSTLIB)
	mkdir -p $(BIN)
	$(CXX) $(CXXFLAGS) $< -o $@ -L$(PKG) -lvc2host -lvc2b200 -lpthread -Wl,-rpath,'$$ORIGIN/..'

# C++ parity test of the Library mirror against the compiled reference (tests/test_gpu_library.py runs it on the GPU box)
$(BIN)/test_library_mirror: tests/cpp/test_library_mirror.cpp include/vc2/*.h $(HOSTLIB)
	mkdir -p $(BIN)
	$(CXX) $(CXXFLAGS) $< -o $@ -L$(PKG) -lvc2host -lvc2b200 -ldl -Wl,-rpath,'$$ORIGIN/..'

# include/vc2/VLC.h against the compiled reference (host only)
$(BIN)/test_vlc_host: tests/cpp/test_vlc_host.cpp include/vc2/VLC.h
	mkdir -p $(BIN)
	$(CXX) $(CXXFLAGS) $< -o $@ -ldl

# host-only check of host/pipeline.h (tests/test_host_helpers.py runs it without a GPU)
$(BIN)/test_pipeline_host: tests/cpp/test_pipeline_host.cpp host/pipeline.h $(LIB)
	mkdir -p $(BIN)
	$(CXX) $(CXXFLAGS) $< -o $@ -L$(PKG) -lvc2b200 -lpthread -Wl,-rpath,'$$ORIGIN/..'

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

$(CSRC)/dwt_tile_fwd.o: $(CSRC)/dwt_tile.cu $(CSRC)/*.cuh include/vc2_cabi.h
	$(NVCC) $(NVFLAGS) -DVC2_DWT_PART=1 -c $< -o $@
$(CSRC)/dwt_tile_inv.o: $(CSRC)/dwt_tile.cu $(CSRC)/*.cuh include/vc2_cabi.h
	$(NVCC) $(NVFLAGS) -DVC2_DWT_PART=2 -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

# A/B probe (tools/pack_probe.c): `make probe-baseline` once on a build the GPU tests passed on, change a kernel,
# `make probe`, then on the GPU box: tools/_probe/pack_probe 32 tools/_probe/baseline.so vc2_reference_b200/libvc2b200.so
# (about 25 s of box time per call: no Python start-up).  tools/_probe/ is git-ignored but travels with gpurun.
probe: $(LIB)
	mkdir -p tools/_probe
	$(CC) -O2 -Wall tools/pack_probe.c -o tools/_probe/pack_probe -ldl
probe-baseline: $(LIB)
	mkdir -p tools/_probe
	cp $(LIB) tools/_probe/baseline.so

clean:
	rm -f $(OBJS) $(LIB) $(HOSTLIB) $(BIN)/EncodeStream $(BIN)/DecodeStream $(BIN)/DecodeFrame
.PHONY: all clean probe probe-baseline
