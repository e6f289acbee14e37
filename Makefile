# Builds libbls_bn254_b200.so for sm_100a without Python (the same command bls_bn254_b200/_native.py and a Rust build.rs run).
NVCC ?= nvcc
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC
LIB := bls_bn254_b200/libbls_bn254_b200.so
SRC := bls_bn254_b200/csrc/capi.cu
HDR := $(wildcard bls_bn254_b200/csrc/*.cuh) include/bls_bn254_b200.h

all: $(LIB)

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVCCFLAGS) -o $@ $(SRC)

# test infrastructure (CPU): the oracle and the host build of the device headers
oracle:
	$(MAKE) -C oracle
hostsim:
	$(MAKE) -C tests/hostsim

# the C++ host-side test program (needs a B200 to pass; exits 77 without one)
tests/cpp/test_api: tests/cpp/test_api.cpp include/bls_bn254_b200.hpp $(LIB)
	$(CXX) -std=c++11 -O1 -Wall -Wextra -Iinclude $< -o $@ -Lbls_bn254_b200 -lbls_bn254_b200 -Wl,-rpath,$(abspath bls_bn254_b200)

clean:
	rm -f $(LIB) tests/cpp/test_api

.PHONY: all oracle hostsim clean
