"""tests/c_caller/v3_caller.c: a C11 program over the V3 half of the ABI, spelled with the safer-ffi generated type names (the surface
the reference's rln/ffi_c_examples use).  CPU: it builds with -Wall -Wextra -Werror and is told there is no GPU; GPU: every step of
the flow (membership, proof, verify / verify_with_roots, record round trip, two-phase proving, secret recovery, stateless object)."""
import os
import subprocess

import pytest

from zerokit_b200 import ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES10 = os.path.join(ROOT, "zerokit_b200", "resources", "tree_depth_10")


def _build(tmp_path):
    ffi.lib()
    exe = str(tmp_path / "v3_caller")
    libdir = os.path.dirname(ffi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_caller", "v3_caller.c"), "-L", libdir, "-lrln_b200", "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_v3_c_caller_links_and_runs(tmp_path):
    import torch
    out = subprocess.run([_build(tmp_path), RES10, "10"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, RLN_B200_WINDOW_BITS="8"))
    assert out.returncode == 0, out.stdout + out.stderr
    assert ("V3-GPU-PATH-OK" if torch.cuda.is_available() else "V3-HOST-PATH-OK") in out.stdout


@pytest.mark.gpu
def test_v3_c_caller_on_gpu(tmp_path):
    out = subprocess.run([_build(tmp_path), RES10, "10"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, RLN_B200_WINDOW_BITS="8"))
    assert out.returncode == 0 and "V3-GPU-PATH-OK" in out.stdout, out.stdout + out.stderr
