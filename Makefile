# Builds libikb200.so (CUDA, sm_100a only) and the CPU oracle.  nvcc cross-compiles without a GPU.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v --expt-relaxed-constexpr
CSRC := ik_b200/csrc
OBJ := build/obj
LIB := ik_b200/libikb200.so

SRCS_CU := $(wildcard $(CSRC)/*.cu)
SRCS_CPP := $(wildcard $(CSRC)/*.cpp)
OBJS := $(patsubst $(CSRC)/%.cu,$(OBJ)/%.cu.o,$(SRCS_CU)) $(patsubst $(CSRC)/%.cpp,$(OBJ)/%.cpp.o,$(SRCS_CPP))
HDRS := $(wildcard $(CSRC)/*.hpp $(CSRC)/*.cuh include/*.h)

all: $(LIB) oracle

$(OBJ)/%.cu.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; exit 1)

$(OBJ)/%.cpp.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p $(OBJ)
	g++ -O2 -std=c++17 -fPIC -Wall -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

oracle:
	$(MAKE) -C oracle -s

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
