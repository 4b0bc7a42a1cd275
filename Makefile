# Builds the B200 (sm_100a) shared library of the rodeo filtering hot path and the CPU oracle port.
#   make            -> rodeo_b200/librodeo_b200.so  +  oracle/_build/librodeo_oracle.so
#   make FAST=1     -> only the FitzHugh-Nagumo instantiations (development builds)
NVCC      ?= nvcc
NVCCFLAGS ?= -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v \
             -diag-suppress 128 $(if $(FAST),-DRODEO_FAST_BUILD,)
CSRC      := rodeo_b200/csrc
OBJDIR    := build/obj$(if $(FAST),_fast,)
TUS       := abi_common abi_dalton abi_solve abi_fenrir abi_hostbuf
OBJS      := $(TUS:%=$(OBJDIR)/%.o)
HDRS      := $(wildcard $(CSRC)/*.cuh) $(CSRC)/rodeo_host.h include/rodeo_b200.h

all: rodeo_b200/librodeo_b200.so oracle

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ > $(OBJDIR)/$*.ptxas.log 2>&1 || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

rodeo_b200/librodeo_b200.so: $(OBJS)
	$(NVCC) -shared -gencode arch=compute_100a,code=sm_100a -o $@ $(OBJS)

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build rodeo_b200/librodeo_b200.so oracle/_build

.PHONY: all oracle clean
