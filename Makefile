# Build of the product library (host C++ + sm_100a CUDA) and of the test oracle.
#   make            -> mujoco_ros_pkgs_b200/libb2mj.so  oracle/liboracle.so
#   make lib / make oracle
ROOT    := $(abspath $(dir $(lastword $(MAKEFILE_LIST))))
CSRC    := $(ROOT)/mujoco_ros_pkgs_b200/csrc
BUILD   ?= $(ROOT)/build
EXTRA   ?=
NVCC    ?= /usr/local/cuda/bin/nvcc
CXX     ?= g++
CXXFLAGS := -std=c++17 -O2 -fPIC -Wall -Wextra -I$(ROOT)/include -I$(CSRC)
NVFLAGS := -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
           -I$(ROOT)/include -I$(CSRC) --expt-relaxed-constexpr -Xptxas -v $(EXTRA)

HOST_SRCS := $(wildcard $(CSRC)/model/*.cpp) $(wildcard $(CSRC)/host/*.cpp)
CUDA_SRCS := $(wildcard $(CSRC)/kernels/*.cu) $(wildcard $(CSRC)/host/*.cu)
HOST_OBJS := $(patsubst $(CSRC)/%.cpp,$(BUILD)/%.o,$(HOST_SRCS))
CUDA_OBJS := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.cu.o,$(CUDA_SRCS)) $(BUILD)/kernels/step_kernel_envmodel.cu.o
LIB ?= $(ROOT)/mujoco_ros_pkgs_b200/libb2mj.so

all: lib oracle
lib: $(LIB)
oracle:
	$(MAKE) -C $(ROOT)/oracle

$(BUILD)/%.o: $(CSRC)/%.cpp $(ROOT)/include/b2mj.h $(wildcard $(CSRC)/*/*.h)
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(BUILD)/%.cu.o: $(CSRC)/%.cu $(ROOT)/include/b2mj.h $(wildcard $(CSRC)/*/*.h) $(wildcard $(CSRC)/*/*.cuh)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) $(NVFLAGS_$(notdir $<)) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)

# second build of the step kernel: per-env model variants (dev_model.h "B2K_PER_ENV_MODEL")
$(BUILD)/kernels/step_kernel_envmodel.cu.o: $(CSRC)/kernels/step_kernel.cu $(ROOT)/include/b2mj.h $(wildcard $(CSRC)/*/*.h) $(wildcard $(CSRC)/*/*.cuh)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -DB2K_PER_ENV_MODEL -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)

# The plugin data paths (robot_hw, sensor readout) are compared BITWISE against oracle/orc_plugins.cpp, which follows the
# reference's C++ line by line (plain IEEE double arithmetic): no FMA contraction in that file.
NVFLAGS_plugins.cu := -fmad=false

$(LIB): $(HOST_OBJS) $(CUDA_OBJS)
	$(NVCC) -shared -gencode arch=compute_100a,code=sm_100a -o $@ $^ -cudart static

clean:
	rm -rf $(BUILD) $(LIB)
	$(MAKE) -C $(ROOT)/oracle clean

.PHONY: all lib oracle clean
