# Convenience targets for callers that do not want Python in the loop.  The Python build (python -m zerokit_b200.build,
# __graft_entry__.build) runs the same nvcc commands.
NVCC     ?= /usr/local/cuda/bin/nvcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default
CSRC     := zerokit_b200/csrc
OBJDIR   := zerokit_b200/lib/obj
LIB      := zerokit_b200/lib/librln_b200.so
UNITS    := rln_host k_poseidon k_prover k_msm_fixed k_msm_var k_verify k_selftest k_records k_witness k_verify_vm
OBJS     := $(UNITS:%=$(OBJDIR)/%.o)
HEADERS  := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.hpp $(CSRC)/*.inc include/*.h)

.PHONY: lib oracle c-example test-cpu clean
lib: $(LIB)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HEADERS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) -ldl

oracle:
	$(MAKE) -C oracle

# the C callers of tests/c_caller, linked against the library (run them on a B200; without a GPU they report the library's error)
c-example: $(LIB)
	gcc -std=c11 -Wall -Wextra -Werror -Iinclude tests/c_caller/basic_caller.c -Lzerokit_b200/lib -lrln_b200 -Wl,-rpath,$(abspath zerokit_b200/lib) -o zerokit_b200/lib/basic_caller
	gcc -std=c11 -Wall -Wextra -Werror -Iinclude tests/c_caller/v3_caller.c -Lzerokit_b200/lib -lrln_b200 -Wl,-rpath,$(abspath zerokit_b200/lib) -o zerokit_b200/lib/v3_caller

test-cpu:
	python -m pytest tests -x -q -m "not gpu"

clean:
	rm -rf $(OBJDIR) $(LIB) oracle/_build tests/host_emul/_build tests/host_fuzz/_build
