# Builds the C-ABI shared library in-tree (the .so travels to the GPU box with the snapshot).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
SRC := $(wildcard neuspeech1_b200/csrc/*.cu)
HDR := $(wildcard neuspeech1_b200/csrc/*.cuh) include/neuspeech_b200.h
OBJ := $(patsubst neuspeech1_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := neuspeech1_b200/lib/libneuspeech_b200.so

all: $(LIB)

build/%.o: neuspeech1_b200/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	@mkdir -p neuspeech1_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)

clean:
	rm -rf build $(LIB)

.PHONY: all clean
