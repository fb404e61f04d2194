# Top-level build: the product library (CUDA, sm_100a only), the checkers under oracle/, and the
# reference-style harness binary ./flash_attention (README.md:83-85 of the reference).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
PKG       := flash_attention_cuda_b200
LIB       := $(PKG)/libflashattn_b200.so
CSRC      := $(PKG)/csrc/fa_api.cu
CHDR      := $(PKG)/csrc/fa_fwd_sm100.cuh $(PKG)/csrc/sm100_ptx.cuh include/flash_attn.h

all: $(LIB) oracle flash_attention

$(LIB): $(CSRC) $(CHDR)
	$(NVCC) $(NVCCFLAGS) -shared $(CSRC) -o $@

oracle:
	$(MAKE) -C oracle

# harness: links the product library; loads the checkers (oracle/) with dlopen at run time
flash_attention: cli/flash_attention_cli.cu $(LIB) include/flash_attn.h
	$(NVCC) $(NVCCFLAGS) -Iinclude cli/flash_attention_cli.cu -o $@ -L$(PKG) -lflashattn_b200 \
	    -Xlinker -rpath -Xlinker '$$ORIGIN/$(PKG)' -ldl

ptxas-info:
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -shared $(CSRC) -o /dev/null

clean:
	rm -f $(LIB) flash_attention
	$(MAKE) -C oracle clean

.PHONY: all oracle clean ptxas-info
