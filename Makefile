# Builds libarkflow_b200.so (sm_100a) in-tree, the C oracle and (in the dev container) nothing from the reference:
# the reference is Rust on top of un-vendored crates and cannot be compiled here.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v --expt-relaxed-constexpr $(EXTRA)
CSRC      := arkflow_b200/csrc
OBJDIR    := build/obj
SRCS      := $(wildcard $(CSRC)/*.cu) $(wildcard $(CSRC)/*.cc) $(wildcard $(CSRC)/*.cpp)
OBJS      := $(patsubst $(CSRC)/%,$(OBJDIR)/%.o,$(SRCS))
LIB       := arkflow_b200/libarkflow_b200.so
HDRS      := $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/arkflow_b200.h

all: $(LIB)

$(OBJDIR)/%.cu.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $@.log || (cat $@.log; exit 1)
	@grep -E "error|warning" $@.log || true

$(OBJDIR)/%.cc.o: $(CSRC)/%.cc $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -x cu -c $< -o $@ 2> $@.log || (cat $@.log; exit 1)
	@grep -E "error|warning" $@.log || true

# pure host code (AVX intrinsics): g++ directly
$(OBJDIR)/%.cpp.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p $(OBJDIR)
	g++ -O3 -std=c++17 -fPIC -Wall -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -Xlinker -soname=libarkflow_b200.so

clean:
	rm -rf build $(LIB)

.PHONY: all clean
