# star-b200 build: hand-written sm_100a CUDA engine + C++ host side -> star_b200/lib/libstar_b200.so, star_b200/bin/STAR
# (the oracle is built separately by oracle/Makefile and oracle/Makefile.ref; the product never links it).
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      := /usr/bin/g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-sign-compare -rdc=false $(NVEXTRA)
CXXFLAGS := -O2 -g -std=c++17 -Wall -Wno-sign-compare -fPIC -pthread
BUILD    := build
ENG_SRC  := $(wildcard star_b200/csrc/engine/*.cu)
ENG_OBJ  := $(patsubst star_b200/csrc/engine/%.cu,$(BUILD)/eng_%.o,$(ENG_SRC))
HOST_SRC := $(wildcard star_b200/csrc/host/*.cpp)
HOST_OBJ := $(patsubst star_b200/csrc/host/%.cpp,$(BUILD)/host_%.o,$(HOST_SRC))
LIB      := star_b200/lib/libstar_b200.so
BIN      := star_b200/bin/STAR

all: $(LIB) $(BIN)

$(BUILD):
	mkdir -p $(BUILD) star_b200/lib star_b200/bin

ENG_HDR  := $(wildcard star_b200/csrc/engine/*.cuh)
$(BUILD)/eng_%.o: star_b200/csrc/engine/%.cu $(ENG_HDR) include/star_b200.h | $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(BUILD)/host_%.o: star_b200/csrc/host/%.cpp star_b200/csrc/host/host.h include/star_b200.h | $(BUILD)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(LIB): $(ENG_OBJ) $(HOST_OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcudart -lpthread -lz

$(BIN): star_b200/csrc/cli_main.cpp $(LIB)
	$(CXX) $(CXXFLAGS) -o $@ $< -Lstar_b200/lib -lstar_b200 -Wl,-rpath,'$$ORIGIN/../lib'

clean:
	rm -rf $(BUILD) $(LIB) $(BIN)
