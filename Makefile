# Builds the product C-ABI library in-tree: fastquick_b200/libfastquick_b200.so
# (hand-written CUDA for sm_100a + C++ host code).  `python -c "import __graft_entry__ as g; g.build()"`
# calls this, then the oracle builds.
NVCC      ?= /usr/local/cuda/bin/nvcc
CSRC      := fastquick_b200/csrc
OBJ       := build/obj
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v --expt-relaxed-constexpr
CXXFLAGS  := -O2 -std=c++17 -fPIC -Wall -I/usr/local/cuda/include
HOST_SRCS := fq_index.cpp fq_synth.cpp fq_capi_host.cpp fq_relayout.cpp fq_hostmath.cpp fq_stats_host.cpp fq_bam.cpp fq_feeder.cpp fq_inflate.cpp fq_deflate.cpp
CU_SRCS   := $(notdir $(wildcard $(CSRC)/*.cu))
OBJS      := $(HOST_SRCS:%.cpp=$(OBJ)/%.o) $(CU_SRCS:%.cu=$(OBJ)/%.cu.o)
LIB       := fastquick_b200/libfastquick_b200.so

CLI       := fastquick_b200/FASTQuick_b200

all: $(LIB) $(CLI)

$(CLI): $(CSRC)/host/fq_cli.cpp $(CSRC)/host/FastQuickB200.cpp $(CSRC)/host/FastQuickB200.h $(LIB)
	g++ $(CXXFLAGS) -o $@ $(CSRC)/host/fq_cli.cpp $(CSRC)/host/FastQuickB200.cpp -Lfastquick_b200 -lfastquick_b200 -Wl,-rpath,'$$ORIGIN' -lz -lpthread

$(OBJ):
	mkdir -p $(OBJ)
$(OBJ)/%.o: $(CSRC)/%.cpp $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/fastquick_b200.h | $(OBJ)
	g++ $(CXXFLAGS) -c $< -o $@
$(OBJ)/%.cu.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/fastquick_b200.h | $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; false)
$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lz -lpthread -lcudart

# development variant with per-kind round counters in search_kernel (tools/quick_bench.py --kstats)
KOBJ := build/kobj
kstats: fastquick_b200/libfastquick_b200_kstats.so
fastquick_b200/libfastquick_b200_kstats.so: $(LIB)
	mkdir -p $(KOBJ)
	$(NVCC) $(NVFLAGS) -DFQB_KSTATS -c $(CSRC)/fq_kernels.cu -o $(KOBJ)/fq_kernels.cu.o 2> $(KOBJ)/fq_kernels.ptxas.log
	$(NVCC) $(ARCH) -shared -o $@ $(filter-out $(OBJ)/fq_kernels.cu.o,$(OBJS)) $(KOBJ)/fq_kernels.cu.o -lz -lpthread -lcudart

clean:
	rm -rf build $(LIB)
.PHONY: all clean kstats
