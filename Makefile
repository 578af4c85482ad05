# Builds libwcmc.so (sm_100a only) in-tree so it travels to the GPU box with the snapshot.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
SRCDIR := wcmc_b200/csrc
SRCS := $(wildcard $(SRCDIR)/*.cu)
OBJS := $(patsubst $(SRCDIR)/%.cu,build/%.o,$(SRCS))
LIB := wcmc_b200/libwcmc.so

all: $(LIB)

build/%.o: $(SRCDIR)/%.cu $(SRCDIR)/common.cuh include/wcmc.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

clean:
	rm -rf build $(LIB)
.PHONY: all clean
