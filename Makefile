# vkhel-b200 build.  Replaces the reference's meson build (meson.build:31-57):
# same product shape -- a static libvkhel_priv.a with every symbol (what the
# reference's tests link, test/meson.build:6,11,16) and a shared libvkhel.so
# that exports vkhel_* only (vkhel.syms) -- built with nvcc for sm_100a.
#
#   make            libraries + oracle + (if /root/reference exists) the
#                   reference's own example and tests, compiled unmodified
#   make DEBUG=1    adds -DVKHEL_DEBUG (reference option debug_logging)
#   make check-host run the host-only reference tests (numbers, ntt)
#   make check      run all reference programs (needs a GPU)

NVCC      ?= /usr/local/cuda/bin/nvcc
# (the image exports CC=/opt/gcc/bin/gcc, a wrapper; use the system gcc)
HOSTCC    ?= /usr/bin/gcc
REF       ?= /root/reference
ARCH      := -gencode arch=compute_100a,code=sm_100a
INC       := -Iinclude -Iinclude/vkhel -Ivkhel_b200/csrc
DEFS      := $(if $(DEBUG),-DVKHEL_DEBUG,)
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall $(INC) $(DEFS)
CFLAGS    := -O2 -fPIC -Wall -Wextra $(INC) $(DEFS)

SRC_DIR := vkhel_b200/csrc
OBJ_DIR := build/obj
LIB_DIR := vkhel_b200/lib
BIN_DIR := build/bin

CU_SRCS := $(SRC_DIR)/device.cu $(SRC_DIR)/vector.cu \
           $(SRC_DIR)/kernels_elem.cu $(SRC_DIR)/kernels_ntt.cu \
           $(SRC_DIR)/tables_device.cu $(SRC_DIR)/probe.cu \
           $(SRC_DIR)/kernels_ntt_cluster.cu $(SRC_DIR)/kernels_ntt_tma.cu \
           $(SRC_DIR)/kernels_ntt_small.cu $(SRC_DIR)/hostcopy.cu
C_SRCS  := $(SRC_DIR)/numbers.c $(SRC_DIR)/ntt_tables.c
OBJS    := $(CU_SRCS:$(SRC_DIR)/%.cu=$(OBJ_DIR)/%.o) \
           $(C_SRCS:$(SRC_DIR)/%.c=$(OBJ_DIR)/%.o)
HDRS    := $(wildcard include/vkhel/*.h include/priv/*.h $(SRC_DIR)/*.cuh)

SHARED := $(LIB_DIR)/libvkhel.so
STATIC := $(LIB_DIR)/libvkhel_priv.a

# link line for a C program against the static library
CUDA_LIBS := -L/usr/local/cuda/lib64 -lcudart_static -lstdc++ -lpthread -ldl -lrt -lm

REF_BINS := $(if $(wildcard $(REF)/test/vector.c), \
	$(BIN_DIR)/ref_example $(BIN_DIR)/ref_test_vector \
	$(BIN_DIR)/ref_test_ntt $(BIN_DIR)/ref_test_numbers,)

.PHONY: all libs oracle refbins examples tools clean check check-host
all: libs oracle refbins examples
examples: $(BIN_DIR)/multi_gpu $(BIN_DIR)/api_loop $(BIN_DIR)/api_product $(BIN_DIR)/api_e2e \
          $(BIN_DIR)/vulkan_dump
libs: $(SHARED) $(STATIC)
refbins: $(REF_BINS)

$(OBJ_DIR)/%.o: $(SRC_DIR)/%.cu $(HDRS)
	@mkdir -p $(OBJ_DIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(OBJ_DIR)/%.o: $(SRC_DIR)/%.c $(HDRS)
	@mkdir -p $(OBJ_DIR)
	$(HOSTCC) $(CFLAGS) -c $< -o $@

$(STATIC): $(OBJS)
	@mkdir -p $(LIB_DIR)
	rm -f $@
	ar rcs $@ $^

$(SHARED): $(OBJS) vkhel.syms
	@mkdir -p $(LIB_DIR)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) \
		-Xlinker --version-script=vkhel.syms -Xlinker -soname=libvkhel.so \
		-Xlinker -z -Xlinker nodelete

# ---- the reference's own programs, compiled from where they lie, unmodified ----
# (include path = include/ and include/vkhel/, as in meson.build:31-32)
$(BIN_DIR)/ref_example: $(REF)/examples/example.c $(SHARED)
	@mkdir -p $(BIN_DIR)
	$(HOSTCC) -O2 $(INC) $< -o $@ -L$(LIB_DIR) -lvkhel -Wl,-rpath,'$$ORIGIN/../../$(LIB_DIR)'

$(BIN_DIR)/ref_test_%: $(REF)/test/%.c $(STATIC)
	@mkdir -p $(BIN_DIR)
	$(HOSTCC) -O2 $(INC) $< -o $@ $(STATIC) $(CUDA_LIBS)

# this repository's own example: limb-sharded transform over every visible GPU
$(BIN_DIR)/multi_gpu: examples/multi_gpu.c $(STATIC)
	@mkdir -p $(BIN_DIR)
	$(HOSTCC) -O2 -Wall $(INC) $< -o $@ $(STATIC) $(CUDA_LIBS)

# the reference API in a loop over separate vectors (recorded transforms)
$(BIN_DIR)/api_loop: examples/api_loop.c $(STATIC)
	@mkdir -p $(BIN_DIR)
	$(HOSTCC) -O2 -Wall $(INC) $< -o $@ $(STATIC) $(CUDA_LIBS)

# the reference's polynomial product sequence (recorded transforms + fused product)
$(BIN_DIR)/api_product: examples/api_product.c $(STATIC)
	@mkdir -p $(BIN_DIR)
	$(HOSTCC) -O2 -Wall $(INC) $< -o $@ $(STATIC) $(CUDA_LIBS)

# the reference's 18 calls end to end from pageable host memory (bench.py's
# e2e_reference_api)
$(BIN_DIR)/api_e2e: examples/api_e2e.c $(STATIC)
	@mkdir -p $(BIN_DIR)
	$(HOSTCC) -O2 -Wall $(INC) $< -o $@ $(STATIC) $(CUDA_LIBS)

# the dump program of tools/vulkan_parity.sh, here against THIS library (the
# same program linked against the meson-built reference shows the shaders)
$(BIN_DIR)/vulkan_dump: tools/vulkan_dump.c $(SHARED)
	@mkdir -p $(BIN_DIR)
	$(HOSTCC) -O2 -Wall $(INC) $< -o $@ -L$(LIB_DIR) -lvkhel -Wl,-rpath,'$$ORIGIN/../../$(LIB_DIR)'

# micro-benchmarks behind the roofline denominators (profiles/*bench*.txt)
tools: $(BIN_DIR)/pipe_bench $(BIN_DIR)/bfly_bench $(BIN_DIR)/exchange_bench \
       $(BIN_DIR)/hostcopy_bench
$(BIN_DIR)/hostcopy_bench: tools/hostcopy_bench.cpp $(OBJ_DIR)/hostcopy.o
	@mkdir -p $(BIN_DIR)
	g++ -O2 $^ -o $@ $(CUDA_LIBS)
$(BIN_DIR)/%_bench: tools/%_bench.cu
	@mkdir -p $(BIN_DIR)
	$(NVCC) $(ARCH) -O3 -o $@ $<

oracle:
	$(MAKE) -C oracle REF=$(REF)

check-host: $(BIN_DIR)/ref_test_numbers $(BIN_DIR)/ref_test_ntt
	$(BIN_DIR)/ref_test_numbers
	$(BIN_DIR)/ref_test_ntt

check: check-host $(BIN_DIR)/ref_test_vector $(BIN_DIR)/ref_example
	$(BIN_DIR)/ref_test_vector
	$(BIN_DIR)/ref_example

clean:
	rm -rf build $(LIB_DIR)
	$(MAKE) -C oracle clean
