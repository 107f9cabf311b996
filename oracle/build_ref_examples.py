"""TEST INFRASTRUCTURE: compiles the reference's own C programs (rln/ffi_c_examples/*.c, written against the
safer-ffi generated `rln.h` of the V3 API) where they lie under /root/reference — never copied — against
include/rln_b200.h and links them with zerokit_b200/lib/librln_b200.so.  Outputs only into oracle/_ref/examples/
(git-ignored, NOT gpurun-ignored: the binaries travel to the GPU box, where /root/reference does not exist, and
tests/test_abi_exports.py::test_reference_c_examples_run_on_gpu executes them there).  These programs are callers of
the product library, i.e. the drop-in check of SURVEY §8b; nothing in the product depends on them."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_EXAMPLES = "/root/reference/rln/ffi_c_examples"
OUT = os.path.join(HERE, "_ref", "examples")
EXAMPLES = ["basic_proof", "multi_message_id", "partial_proof", "recover_secret", "stateless", "type_serialization"]


def build():
    """no-op (returns None) when the reference checkout is absent — the GPU box uses the prebuilt files"""
    if not os.path.isdir(REF_EXAMPLES):
        return None
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "rln.h"), "w") as f:      # the only shim: the header name the examples include
        f.write('#include "rln_b200.h"\n')
    libdir = os.path.join(ROOT, "zerokit_b200", "lib")
    for ex in EXAMPLES:
        exe = os.path.join(OUT, ex)
        # rpath relative to the binary: the snapshot lives at another absolute path on the GPU box
        subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", OUT, "-I", os.path.join(ROOT, "include"),
                               "-I", REF_EXAMPLES, os.path.join(REF_EXAMPLES, ex + ".c"), "-L", libdir, "-lrln_b200",
                               "-Wl,-rpath,$ORIGIN/../../../zerokit_b200/lib", "-o", exe])
    return OUT


if __name__ == "__main__":
    print(build())
