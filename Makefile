# Top-level build: the CUDA C-ABI library, the host program, and the test oracle.
#
#   make            libgcn10cuda.so + host program (gcn10) + oracle
#   make cuda       gcn10_b200/libgcn10cuda.so      (nvcc, sm_100a only)
#   make host       gcn10_b200/host/gcn10           (gcc, links libgcn10cuda.so + zlib)
#   make oracle     oracle/libcn_oracle.so and, when /root/reference exists, oracle/_ref/
#   make integration  oracle/_ref/libgcn10_gpu_ref.so: integration/cn_gpu.c + the reference's raster.c + libgcn10cuda
NVCC      ?= nvcc
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall -Xptxas -v

CUDA_SO := gcn10_b200/libgcn10cuda.so
CUDA_SRC := gcn10_b200/csrc/gcn10_cuda.cu
CUDA_HDR := $(wildcard gcn10_b200/csrc/*.cuh) $(wildcard gcn10_b200/csrc/*.h) include/gcn10_cuda.h

all: cuda host oracle integration

cuda: $(CUDA_SO)

$(CUDA_SO): $(CUDA_SRC) $(CUDA_HDR)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CUDA_SRC) > gcn10_b200/csrc/ptxas.log 2>&1 || (cat gcn10_b200/csrc/ptxas.log; exit 1)

host: cuda
	@if [ -f gcn10_b200/host/Makefile ]; then $(MAKE) -s -C gcn10_b200/host; fi

oracle:
	$(MAKE) -s -C oracle all

# the reference-side binding (integration/cn_gpu.c) built against the reference's own headers and raster.c
integration: cuda
	$(MAKE) -s -C integration all

clean:
	rm -f $(CUDA_SO) gcn10_b200/csrc/ptxas.log
	$(MAKE) -s -C oracle clean
	@if [ -f gcn10_b200/host/Makefile ]; then $(MAKE) -s -C gcn10_b200/host clean; fi

.PHONY: all cuda host oracle integration clean
