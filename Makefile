# Builds libargus_b200.so (sm_100a only) in-tree, plus the C oracle helpers.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
EXTRA     ?=
BUILD     ?= build
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr $(EXTRA)
CSRC      := argus_b200/csrc
SOURCES   := $(wildcard $(CSRC)/*.cu)
HEADERS   := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/argus_b200.h
OBJECTS   := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.o,$(SOURCES))
LIB       ?= argus_b200/libargus_b200.so

all: $(LIB)

$(BUILD)/%.o: $(CSRC)/%.cu $(HEADERS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJECTS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJECTS) -lcudart_static -lpthread -ldl -lrt

ptxas-info:
	@mkdir -p build
	for f in $(SOURCES); do $(NVCC) $(NVFLAGS) -Xptxas -v -c $$f -o /dev/null 2>&1 | grep -E "Compiling|registers|spill" ; done

clean:
	rm -rf build $(LIB)

.PHONY: all clean ptxas-info
