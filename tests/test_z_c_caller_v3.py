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


@pytest.mark.gpu
def test_tree_out_of_bounds_like_reference():
    """rln/tests/ffi.rs:1070-1173 (test_rln_out_of_bounds_ffi): every index at or past the capacity, overflowing ranges, delete indices
    after the start of an override, too many initial leaves and a full tree are refused with an error result — never a crash"""
    os.environ.setdefault("RLN_B200_WINDOW_BITS", "8")
    import zerokit_b200 as z
    rln = z.RLN.new(10)
    cap = 1 << 10
    for bad in (cap + 10, cap):
        for op in (lambda: rln.set_leaf(bad, 123), lambda: rln.get_merkle_proof(bad), lambda: rln.get_leaf(bad), lambda: rln.delete_leaf(bad)):
            with pytest.raises(z.RLNError):
                op()
    with pytest.raises(z.RLNError):
        rln.set_leaves_from(cap + 10, [1, 2])
    with pytest.raises(z.RLNError):
        rln.atomic_operation(cap + 10, [1], [cap + 10])
    with pytest.raises(z.RLNError):
        rln.atomic_operation(0, [1], [cap + 10])
    with pytest.raises(z.RLNError):
        rln.atomic_operation(0, [1, 2], [1])              # delete index after `start`
    with pytest.raises(z.RLNError):
        rln.atomic_operation(2 ** 64 - 1, [1], [0])       # start + len overflows usize
    empty_root = rln.get_root()
    assert rln.leaves_set() == 0                          # nothing above changed the tree
    rln.set_tree(4)
    with pytest.raises(z.RLNError):
        rln.init_tree_with_leaves(list(range(1, 18)))     # 17 leaves into 16 slots
    for _ in range(16):
        rln.set_next_leaf(3)
    with pytest.raises(z.RLNError):
        rln.set_next_leaf(3)                              # tree is full
    assert rln.leaves_set() == 16
    rln.set_tree(10)
    assert rln.get_root() == empty_root
