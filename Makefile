# Builds the B200 (sm_100a) shared library of the rodeo filtering hot path and the CPU oracle port.
#   make            -> rodeo_b200/librodeo_b200.so  +  oracle/_build/librodeo_oracle.so
#   make FAST=1     -> only the FitzHugh-Nagumo instantiations (development builds)
#   make FAST=1 VARIANT=r128 EXTRA=-DRODEO_DALTON_MAXNREG=128 OUT=build/variants/r128.so   (tuning experiments;
#        select at run time with RODEO_B200_LIB=build/variants/r128.so)
NVCC      ?= nvcc
NVCCFLAGS ?= -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v \
             -diag-suppress 128 $(if $(FAST),-DRODEO_FAST_BUILD,) $(EXTRA)
CSRC      := rodeo_b200/csrc
OBJDIR    := build/obj$(if $(FAST),_fast,)$(if $(VARIANT),_$(VARIANT),)
OUT       ?= rodeo_b200/librodeo_b200.so
TUS       := abi_common abi_dalton abi_solve abi_solve_sim abi_fenrir abi_hostbuf abi_nvrtc embedded_headers \
             abi_dalton_f32 abi_solve_f32 abi_solve_sim_f32 abi_fenrir_f32 abi_dalton_solve_mv abi_dalton_solve_sim abi_solve_sqrt abi_solve_sqrt_f32 abi_fenrir_solve abi_kalmantv abi_kalmantv_sqrt abi_magi abi_mcmc abi_peer
OBJS      := $(TUS:%=$(OBJDIR)/%.o)
HDRS      := $(wildcard $(CSRC)/*.cuh) $(CSRC)/rodeo_host.h include/rodeo_b200.h

all: $(OUT) oracle

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ > $(OBJDIR)/$*.ptxas.log 2>&1 || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

# translation-unit shims that #include another .cu with a different arithmetic type / op
$(OBJDIR)/abi_dalton_f32.o: $(CSRC)/abi_dalton.cu
$(OBJDIR)/abi_solve_f32.o: $(CSRC)/abi_solve.cu
$(OBJDIR)/abi_solve_sim_f32.o: $(CSRC)/abi_solve_sim.cu
$(OBJDIR)/abi_fenrir_f32.o: $(CSRC)/abi_fenrir.cu
$(OBJDIR)/abi_solve_sqrt_f32.o: $(CSRC)/abi_solve_sqrt.cu
$(OBJDIR)/abi_dalton_solve_mv.o $(OBJDIR)/abi_dalton_solve_sim.o: $(CSRC)/abi_dalton_solve.cu

# the three device headers as string literals, for NVRTC (abi_nvrtc.cu)
$(OBJDIR)/embedded_headers.cu: $(wildcard $(CSRC)/*.cuh) tools/embed_headers.py
	@mkdir -p $(OBJDIR)
	python tools/embed_headers.py $(CSRC) $@

$(OBJDIR)/embedded_headers.o: $(OBJDIR)/embedded_headers.cu
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ > $(OBJDIR)/embedded_headers.ptxas.log 2>&1 || (cat $(OBJDIR)/embedded_headers.ptxas.log; exit 1)

$(OUT): $(OBJS)
	@mkdir -p $(dir $(OUT))
	$(NVCC) -shared -gencode arch=compute_100a,code=sm_100a -o $@ $(OBJS) -ldl

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build rodeo_b200/librodeo_b200.so oracle/_build

.PHONY: all oracle clean
