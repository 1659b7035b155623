# Builds libmpc_b200.so (CUDA kernels + C ABI, sm_100a only), the host CLI `compressor` (drop-in for the
# reference binary that bin/run invokes) and the CPU oracle used by the tests.
PKG      := cal_22-mpc_b200
CSRC     := $(PKG)/csrc
HOST     := $(PKG)/host
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      := /usr/bin/g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -I$(CSRC) --expt-relaxed-constexpr
CXXFLAGS := -O3 -std=c++17 -fPIC -Iinclude -I$(CSRC) -Wall

LIB      := $(PKG)/libmpc_b200.so
SPEC_CFGS := P6 F4 Z1 E5 S32 S64
SPEC_SRCS := $(foreach c,$(SPEC_CFGS),$(CSRC)/spec/spec_$(c).cu)
CU_SRCS  := $(CSRC)/mpc_capi.cu $(CSRC)/mpc_generic.cu $(CSRC)/mpc_synth.cu $(CSRC)/mpc_variants.cu $(CSRC)/mpc_sc2.cu $(CSRC)/mpc_pattern.cu $(CSRC)/mpc_spec_registry.cu $(CSRC)/mpc_spec_list.cu $(SPEC_SRCS)
CU_OBJS  := $(CU_SRCS:.cu=.o)
CC_OBJS  := $(CSRC)/mpc_config.o $(CSRC)/mpc_specgen.o $(CSRC)/mpc_jit.o

all: $(LIB) compressor oracle tools/int_peak

# integer-pipe issue-rate micro-benchmark (roofline B); bench.py runs it on the GPU box
tools/int_peak: tools/int_peak.cu
	$(NVCC) $(ARCH) -O3 -lineinfo -o $@ $<

# config compiler (csrc/mpc_specgen.cpp): one specialised schedule per shipped config; generated sources are committed
tools/specgen: tools/specgen_main.cpp $(CSRC)/mpc_specgen.cpp $(CSRC)/mpc_specgen.h $(CSRC)/mpc_layout.h $(CSRC)/mpc_config.cpp include/mpc_capi.h
	$(CXX) -O1 -std=c++17 -Iinclude -I$(CSRC) tools/specgen_main.cpp $(CSRC)/mpc_specgen.cpp $(CSRC)/mpc_config.cpp -o $@
$(CSRC)/spec/.stamp: tools/specgen $(foreach c,$(SPEC_CFGS),configs/$(c).json)
	@mkdir -p $(CSRC)/spec
	tools/specgen $(CSRC)/spec $(foreach c,$(SPEC_CFGS),configs/$(c).json)
	@touch $@
# the empty recipe (";") makes make re-read the time stamps of the generated sources after the generator ran
$(SPEC_SRCS) $(CSRC)/spec/spec_list.inc: $(CSRC)/spec/.stamp ;

$(CSRC)/mpc_spec_list.o: $(CSRC)/spec/spec_list.inc

$(CSRC)/spec/%.o: $(CSRC)/spec/%.cu $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(@:.o=.ptxas.txt) || (cat $(@:.o=.ptxas.txt); false)

$(CSRC)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)
	$(NVCC) $(NVFLAGS) -c $< -o $@

# headers the run-time (NVRTC) build of the specialised kernel compiles against, embedded as raw string literals
EMBED_HDRS := mpc_spec.cuh mpc_device.cuh mpc_layout.h
$(CSRC)/mpc_embedded_headers.inc: $(foreach h,$(EMBED_HDRS),$(CSRC)/$(h))
	@rm -f $@; for h in $(EMBED_HDRS); do \
	  printf 'static const char kHdr_%s[] = R"MPCHDR(' "$$(echo $$h | tr . _)" >> $@; cat $(CSRC)/$$h >> $@; printf ')MPCHDR";\n' >> $@; done
$(CSRC)/mpc_jit.o: $(CSRC)/mpc_embedded_headers.inc

$(CSRC)/%.o: $(CSRC)/%.cpp $(wildcard $(CSRC)/*.h include/*.h)
	$(CXX) $(CXXFLAGS) -I/usr/local/cuda/include -c $< -o $@

$(LIB): $(CU_OBJS) $(CC_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart shared -lnvrtc -ldl -Xlinker --no-undefined -Xlinker -rpath=/usr/local/cuda/lib64

HOST_SRCS := $(wildcard $(HOST)/*.cpp $(HOST)/compressor/*.cpp $(HOST)/loader/*.cpp)
compressor: bin/compressor
bin/compressor: $(HOST_SRCS) $(wildcard $(HOST)/*.h $(HOST)/compressor/*.h $(HOST)/loader/*.h) $(LIB)
	@mkdir -p bin; if [ -n "$(HOST_SRCS)" ]; then \
	  $(CXX) $(CXXFLAGS) -I$(HOST) -I/usr/local/cuda/include $(HOST_SRCS) -o bin/compressor -L$(PKG) -lmpc_b200 \
	    -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$$ORIGIN/../$(PKG)' -Wl,-rpath,/usr/local/cuda/lib64 -lpthread; \
	fi

oracle:
	$(MAKE) -C oracle

# SASS of the specialised kernels' hot loops for profiles/ (one listing per shipped config)
sass: $(LIB)
	python3 tools/sass_listing.py

clean:
	rm -f $(CSRC)/*.o $(CSRC)/spec/*.o $(LIB) bin/compressor tools/specgen
	$(MAKE) -C oracle clean

.PHONY: all oracle clean sass compressor
