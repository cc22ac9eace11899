# Builds libjstsp_b200.so (sm_100a only) in-tree, plus the C oracle helpers.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall -Xptxas -v --expt-relaxed-constexpr
CSRC      := jstsp19_b200/csrc
OBJDIR    := build/obj
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
LIB       := jstsp19_b200/libjstsp_b200.so

all: $(LIB)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/jstsp_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

clean:
	rm -rf build/obj build/racecheck build/debug $(LIB)

# sanitizer build: the streamed-ring kernels release their slots with a CTA barrier (stream_core.cuh), selected by JSTSP_LIB at load time
RCDIR := build/racecheck
RCOBJS := $(patsubst $(CSRC)/%.cu,$(RCDIR)/%.o,$(SRCS))
$(RCDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/jstsp_b200.h
	@mkdir -p $(RCDIR)
	$(NVCC) $(NVCCFLAGS) -DJSTSP_PIPE_CTA_SYNC -c $< -o $@ 2> $(RCDIR)/$*.ptxas.log || (cat $(RCDIR)/$*.ptxas.log; exit 1)
# instrumented build of the persistent kernel (phase stamps / beacons / dumps of admm_mega.cuh) for tools/mega_*.py
DBGDIR := build/debug
DBGOBJS := $(patsubst $(CSRC)/%.cu,$(DBGDIR)/%.o,$(SRCS))
$(DBGDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/jstsp_b200.h
	@mkdir -p $(DBGDIR)
	$(NVCC) $(NVCCFLAGS) -DJSTSP_MEGA_DEBUG=1 -c $< -o $@ 2> $(DBGDIR)/$*.ptxas.log || (cat $(DBGDIR)/$*.ptxas.log; exit 1)
debug-lib: $(DBGOBJS)
	$(NVCC) $(ARCH) -shared -o $(DBGDIR)/libjstsp_b200.so $(DBGOBJS) -lcudart
racecheck-lib: $(RCOBJS)
	$(NVCC) $(ARCH) -shared -o $(RCDIR)/libjstsp_b200.so $(RCOBJS) -lcudart

.PHONY: all clean racecheck-lib debug-lib

# ---- MEX gateways against the in-repo mex.h shim (unit-test build; a MATLAB user runs `mex -R2018a`) ----
MEXDIR  := jstsp19_b200/mex
MEXSRC  := $(filter-out $(MEXDIR)/shim/%,$(wildcard $(MEXDIR)/*.c))
MEXOUT  := $(patsubst $(MEXDIR)/%.c,$(MEXDIR)/build/%.so,$(MEXSRC))

mex-shim: $(LIB) $(MEXOUT)

$(MEXDIR)/build/%.so: $(MEXDIR)/%.c $(MEXDIR)/gateway_common.h $(MEXDIR)/shim/mex.h $(MEXDIR)/shim/mex_shim.c include/jstsp_b200.h
	@mkdir -p $(MEXDIR)/build
	gcc -O2 -Wall -Wno-misleading-indentation -Wno-unused-function -shared -fPIC -I$(MEXDIR)/shim -o $@ $< $(MEXDIR)/shim/mex_shim.c -Ljstsp19_b200 -ljstsp_b200 -Wl,-rpath,'$$ORIGIN/../..' -lm

.PHONY: mex-shim
