# Builds libb2dgpu.so (sm_100a) in-tree.  `make` here is what __graft_entry__.build() runs.
NVCC      ?= /usr/local/cuda/bin/nvcc
CSRC      := blend2d_b200/csrc
OUT       := blend2d_b200/libb2dgpu.so
NVFLAGS   := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
             -Xcompiler -fPIC,-fvisibility=hidden,-O2 -Xptxas -v --expt-relaxed-constexpr $(EXTRA)
HOSTSRC   := $(wildcard $(CSRC)/host/*.cpp)
DEVHDR    := $(wildcard $(CSRC)/*.cuh) $(CSRC)/kernels.h include/b2dgpu.h

all: $(OUT)

$(CSRC)/kernels.o: $(CSRC)/kernels.cu $(DEVHDR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(CSRC)/kernels.ptxas.log || (cat $(CSRC)/kernels.ptxas.log; false)

$(CSRC)/stream.o: $(CSRC)/stream.cu $(DEVHDR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(CSRC)/stream.ptxas.log || (cat $(CSRC)/stream.ptxas.log; false)

$(CSRC)/runtime.o: $(CSRC)/runtime.cu $(DEVHDR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(CSRC)/host/%.o: $(CSRC)/host/%.cpp $(wildcard $(CSRC)/host/*.h) include/b2dgpu.h
	g++ -std=c++17 -O2 -fPIC -fvisibility=hidden -ffp-contract=off -Iinclude -c $< -o $@

$(OUT): $(CSRC)/kernels.o $(CSRC)/stream.o $(CSRC)/runtime.o $(HOSTSRC:.cpp=.o)
	$(NVCC) -shared -o $@ $^ -cudart static -Xlinker --no-undefined -lpthread -ldl -lrt

clean:
	rm -f $(CSRC)/*.o $(CSRC)/host/*.o $(OUT) $(CSRC)/*.log
.PHONY: all clean
