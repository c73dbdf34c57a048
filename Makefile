# Builds everything in-tree (the .so files are git-ignored but travel to the GPU box).
#   make gpu     dsf2flac_b200/libdsdgpu.so    CUDA kernels + the C ABI of include/dsdgpu.h (sm_100a)
#   make host    dsf2flac_b200/libdsdhost.so   drop-in DsdDecimator + a C-ABI driver for the tests
#   make oracle  oracle/_build, oracle/_ref    the CPU checkers (test infrastructure only)
#   make tests   tests/cpp/_build/*            lane emulator, CPU stand-in of the C ABI for host-logic tests
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
PKG := dsf2flac_b200
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall \
           -I include -I $(PKG)/csrc
HOSTSRC := $(PKG)/host/dsd_decimator.cpp $(PKG)/host/dsd_container.cpp $(PKG)/host/dsdhost_capi.cpp $(PKG)/host/dsd_synth.cpp
HOSTHDR := $(wildcard $(PKG)/host/*.h) $(wildcard $(PKG)/host/standalone/*.h) $(wildcard $(PKG)/host/*.inc) include/dsdgpu.h
# host/standalone holds the Boost/id3lib-free mirror of the reference's dsd_sample_reader.h; inside dsf2flac (and in
# oracle/_ref/libdropin_ref*.so, see oracle/Makefile) the same dsd_decimator.{h,cpp} compile against the reference's own header
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall -ffp-contract=off -I include -I $(PKG)/host -I $(PKG)/host/standalone

all: gpu host oracle tests

gpu: $(PKG)/libdsdgpu.so
$(PKG)/libdsdgpu.so: $(wildcard $(PKG)/csrc/*.cpp) $(wildcard $(PKG)/csrc/*.cu) $(wildcard $(PKG)/csrc/*.h) $(wildcard $(PKG)/csrc/*.cuh) include/dsdgpu.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(wildcard $(PKG)/csrc/*.cu) $(wildcard $(PKG)/csrc/*.cpp)

host: $(PKG)/libdsdhost.so
$(PKG)/libdsdhost.so: $(HOSTSRC) $(HOSTHDR) $(PKG)/libdsdgpu.so
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOSTSRC) -L$(PKG) -ldsdgpu -Wl,-rpath,'$$ORIGIN'

oracle: gpu
	$(MAKE) -C oracle all

tests: tests/cpp/_build/emu_ring tests/cpp/_build/libdsdhost_cpu.so tests/cpp/_build/libdsdsynth.so
tests/cpp/_build/emu_ring: tests/cpp/emu_ring.cpp $(PKG)/csrc/ring_core.h
	mkdir -p tests/cpp/_build
	$(CXX) -O2 -std=c++17 -I $(PKG)/csrc -o $@ tests/cpp/emu_ring.cpp
# the synthetic-input generator on its own (bench.py's reference arm must not map the product libraries)
tests/cpp/_build/libdsdsynth.so: $(PKG)/host/dsd_synth.cpp
	mkdir -p tests/cpp/_build
	$(CXX) -O2 -std=c++17 -fPIC -shared -o $@ $(PKG)/host/dsd_synth.cpp
# the same host sources linked against a CPU stand-in of the C ABI (tests only, never shipped)
tests/cpp/_build/libdsdhost_cpu.so: $(HOSTSRC) $(HOSTHDR) tests/cpp/standin_dsdgpu.cpp $(PKG)/csrc/shard_plan.cpp $(PKG)/csrc/batch_run.cpp
	mkdir -p tests/cpp/_build
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOSTSRC) tests/cpp/standin_dsdgpu.cpp $(PKG)/csrc/shard_plan.cpp $(PKG)/csrc/batch_run.cpp -lpthread

clean:
	rm -rf $(PKG)/*.so tests/cpp/_build
	$(MAKE) -C oracle clean

.PHONY: all gpu host oracle tests clean
