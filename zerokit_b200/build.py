"""Builds zerokit_b200/lib/librln_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

One object per translation unit, compiled in parallel; an object is rebuilt when its source or any
header in csrc/ or include/ is newer.  Usage: python -m zerokit_b200.build [--force]
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# A/B builds: RLN_B200_BUILD_VARIANT=name compiles the units listed in RLN_B200_BUILD_VARIANT_UNITS (comma separated) with the extra
# flags of RLN_B200_BUILD_VARIANT_FLAGS into lib/obj_<name>/ and links lib/librln_b200_<name>.so with the other units' plain objects
VARIANT = os.environ.get("RLN_B200_BUILD_VARIANT", "")
VARIANT_UNITS = [u for u in os.environ.get("RLN_B200_BUILD_VARIANT_UNITS", "").split(",") if u]
VARIANT_FLAGS = os.environ.get("RLN_B200_BUILD_VARIANT_FLAGS", "").split()
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "librln_b200.so" if not VARIANT else f"librln_b200_{VARIANT}.so")
UNITS = ["rln_host.cu", "k_poseidon.cu", "k_prover.cu", "k_msm_fixed.cu", "k_msm_var.cu", "k_verify.cu", "k_selftest.cu", "k_records.cu", "k_witness.cu", "k_verify_vm.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=default"]


def _headers_mtime():
    m = 0.0
    for d in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(d):
            if f.endswith((".cuh", ".hpp", ".h", ".inc")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def _obj(unit):
    d = OBJ + "_" + VARIANT if (VARIANT and unit in VARIANT_UNITS) else OBJ
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, unit.replace(".cu", ".o"))


def _compile(unit, force, hm):
    src = os.path.join(CSRC, unit)
    obj = _obj(unit)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hm):
        return unit, False, ""
    extra = VARIANT_FLAGS if (VARIANT and unit in VARIANT_UNITS) else []
    r = subprocess.run([NVCC] + FLAGS + extra + ["-c", src, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {unit}:\n{r.stdout}\n{r.stderr}")
    return unit, True, r.stderr


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    hm = _headers_mtime()
    rebuilt = False
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        for unit, did, _ in ex.map(lambda u: _compile(u, force, hm), UNITS):
            rebuilt = rebuilt or did
            if verbose and did:
                print(f"[build] compiled {unit}", flush=True)
    objs = [_obj(u) for u in UNITS]
    if rebuilt or not os.path.exists(LIB):
        r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-ldl"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[build] linked {LIB}", flush=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
