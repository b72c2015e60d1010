# Builds the product library (CUDA, sm_100a only) and, for tests/hostsim only, the host-emulation
# build of the same device sources. The oracle has its own recipe in oracle/Makefile.
NVCC ?= nvcc
HOSTCXX := /usr/bin/g++
CSRC := latticednaorigami_b200/csrc
OUT := latticednaorigami_b200
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ccbin $(HOSTCXX) -Xcompiler -fPIC
CXXFLAGS := -std=c++17 -O2 -fPIC -Wall -Wno-unused-function -Wno-unknown-pragmas
HOST_SRCS := $(CSRC)/ldo_host.cpp $(CSRC)/ldo_sim.cpp
HDRS := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.hpp include/*.h)

all: $(OUT)/libldo_b200.so $(OUT)/latticeDNAOrigami_b200

$(OUT)/build/ldo_engine.o: $(CSRC)/ldo_engine.cu $(HDRS)
	mkdir -p $(OUT)/build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OUT)/build/%.o: $(CSRC)/%.cpp $(HDRS)
	mkdir -p $(OUT)/build
	$(HOSTCXX) $(CXXFLAGS) -c $< -o $@

$(OUT)/libldo_b200.so: $(OUT)/build/ldo_engine.o $(OUT)/build/ldo_host.o $(OUT)/build/ldo_sim.o
	$(NVCC) -shared -ccbin $(HOSTCXX) -o $@ $^ -lcudart -ldl

$(OUT)/latticeDNAOrigami_b200: $(CSRC)/ldo_main.cpp $(OUT)/libldo_b200.so
	$(HOSTCXX) $(CXXFLAGS) -o $@ $< -L$(OUT) -lldo_b200 -lpthread -Wl,-rpath,'$$ORIGIN'

# Device-logic emulation on the host (one emulated lane per replica); test infrastructure only.
hostsim: tests/hostsim/libldo_hostsim.so
tests/hostsim/libldo_hostsim.so: $(CSRC)/ldo_engine.cu $(HOST_SRCS) $(HDRS)
	mkdir -p tests/hostsim
	$(HOSTCXX) $(CXXFLAGS) -O1 -g -DLDO_HOSTSIM -shared -o $@ -x c++ $(CSRC)/ldo_engine.cu $(HOST_SRCS)

# A/B variants for profiling: make variant NAME=generic DEFS="-DLDO_GENERIC_ACCESS" -> ab/lib_generic.so
# (selected at run time with LDO_B200_LIB=ab/lib_generic.so)
variant:
	mkdir -p ab
	$(NVCC) $(NVFLAGS) $(DEFS) -c $(CSRC)/ldo_engine.cu -o ab/ldo_engine_$(NAME).o
	$(NVCC) -shared -ccbin $(HOSTCXX) -o ab/lib_$(NAME).so ab/ldo_engine_$(NAME).o $(OUT)/build/ldo_host.o $(OUT)/build/ldo_sim.o -lcudart -ldl

clean:
	rm -rf $(OUT)/build $(OUT)/libldo_b200.so $(OUT)/latticeDNAOrigami_b200 tests/hostsim/libldo_hostsim.so

.PHONY: all hostsim clean variant
