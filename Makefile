# libksg_b200.so (the C ABI of include/ksg_b200.h, sm_100a only) and the C-ABI example.
# The Python package builds the same library through syntropy_b200/_build.py.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
LIB  := syntropy_b200/lib/libksg_b200.so

all: $(LIB) examples/ksg_mi_cabi

$(LIB): $(wildcard syntropy_b200/csrc/*) include/ksg_b200.h
	python -m syntropy_b200._build

examples/ksg_mi_cabi: examples/ksg_mi_cabi.cu include/ksg_b200.h $(LIB)
	$(NVCC) $(ARCH) -O2 -std=c++17 -I include -o $@ $< -L syntropy_b200/lib -lksg_b200 \
	    -Xlinker -rpath -Xlinker $(CURDIR)/syntropy_b200/lib

clean:
	rm -rf syntropy_b200/build syntropy_b200/lib examples/ksg_mi_cabi

.PHONY: all clean
