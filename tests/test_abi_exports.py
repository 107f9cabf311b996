"""CPU: the C-ABI library loads without a GPU, exports every symbol include/rln_b200.h declares, and
refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import subprocess

import pytest

from zerokit_b200 import ffi


def test_exports_match_header():
    L = ffi.lib()
    out = subprocess.check_output(["nm", "-D", "--defined-only", ffi.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if l.strip()}
    declared = ffi.declared_symbols()
    assert len(declared) > 80
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    unbound = [s for s in declared if s not in L._signatures]
    assert not unbound, unbound


def test_host_only_helpers_work_without_gpu(goldens):
    import zerokit_b200 as z
    for msg, v in goldens["derived"]["hash_to_field"].items():
        assert z.hash_to_field_le(msg.encode()) == int(v)
        assert z.hash_to_field_be(msg.encode()) == int(v)
    w = z.RLNWitnessInput.new_single(5, 10, 3, [1, 2], [0, 1], 7, 9)
    b = w.to_bytes_le()
    assert len(b) == 1 + 32 * 7 + 16 + 2 and z.RLNWitnessInput.from_bytes_le(b).to_bytes_le() == b
    with pytest.raises(z.RLNError, match="User message limit cannot be zero"):
        z.RLNWitnessInput.new_single(5, 0, 0, [1], [0], 7, 9)
    with pytest.raises(z.RLNError, match=r"Message id \(10\) is not within user_message_limit \(10\)"):
        z.RLNWitnessInput.new_single(5, 10, 10, [1], [0], 7, 9)
    with pytest.raises(z.RLNError, match="Merkle proof length mismatch: expected 2, got 1"):
        z.RLNWitnessInput.new_single(5, 10, 1, [1, 2], [0], 7, 9)
    with pytest.raises(z.RLNError, match="Expected to read"):
        z.RLNWitnessInput.from_bytes_le(b[:-1])


def test_multi_message_id_witness_codec(goldens):
    """MultiV1 wire layout (rln/src/protocol/mode.rs:26-74, witness.rs:115-176,400-413) — host-only"""
    import sys, os
    import zerokit_b200 as z
    from pyref import poseidon as P
    k = goldens["derived"]["kat_proof_multi_d20"]
    pe = [P.poseidon([i + 7]) for i in range(20)]
    idx = [(5 * i + 1) % 2 for i in range(20)]
    w = z.RLNWitnessInput.new_multi(424242, 50, [3, 7, 11, 0], pe, idx, 1234567, 89, [True, False, True, False])
    assert w.to_bytes_le().hex() == k["witness_le_hex"]
    assert z.RLNWitnessInput.from_bytes_le(bytes.fromhex(k["witness_le_hex"])).to_bytes_le().hex() == k["witness_le_hex"]
    with pytest.raises(z.RLNError, match="At least one selector_used value must be true"):
        z.RLNWitnessInput.new_multi(1, 50, [3, 7], pe, idx, 1, 1, [False, False])
    with pytest.raises(z.RLNError, match="Duplicate message ID"):
        z.RLNWitnessInput.new_multi(1, 50, [3, 3], pe, idx, 1, 1, [True, True])
    z.RLNWitnessInput.new_multi(1, 50, [3, 3], pe, idx, 1, 1, [True, False])   # duplicates only count among used slots
    with pytest.raises(z.RLNError, match=r"Message id \(50\) is not within user_message_limit \(50\)"):
        z.RLNWitnessInput.new_multi(1, 50, [3, 50], pe, idx, 1, 1, [True, True])
    z.RLNWitnessInput.new_multi(1, 50, [3, 50], pe, idx, 1, 1, [True, False])  # unused slots are not range-checked
    with pytest.raises(z.RLNError, match="The field message_ids has length 2, but the field selector_used has length 1"):
        z.RLNWitnessInput.new_multi(1, 50, [3, 4], pe, idx, 1, 1, [True])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import zerokit_b200 as z
    with pytest.raises(z.RLNError, match="no usable CUDA device"):
        z.RLN.new(20)
    with pytest.raises(z.RLNError):
        z.poseidon_hash([1, 2])


def _build_c_caller(tmp_path):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "basic_caller")
    libdir = os.path.dirname(ffi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "c_caller", "basic_caller.c"), "-L", libdir, "-lrln_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_plain_c_caller_links_and_runs(tmp_path):
    """include/rln_b200.h is valid C11 and a C program written against it (the flow of rln/ffi_c_examples/basic_proof.c)
    links with -lrln_b200; without a GPU it must hear 'no usable CUDA device' and still get the host-only codecs"""
    import torch
    ffi.lib()
    exe = _build_c_caller(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert ("GPU-PATH-OK" if torch.cuda.is_available() else "HOST-PATH-OK") in out.stdout


@pytest.mark.gpu
def test_plain_c_caller_on_gpu(tmp_path):
    ffi.lib()
    out = subprocess.run([_build_c_caller(tmp_path)], capture_output=True, text=True, env={"RLN_B200_WINDOW_BITS": "8", "PATH": "/usr/bin:/bin"})
    assert out.returncode == 0 and "GPU-PATH-OK" in out.stdout, out.stderr + out.stdout


REF_EXAMPLES = "/root/reference/rln/ffi_c_examples"


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("example", ["basic_proof", "multi_message_id", "partial_proof", "recover_secret", "stateless", "type_serialization"])
def test_reference_c_examples_build_unmodified(tmp_path, example):
    """Source-level drop-in: the reference's own C examples (written against the safer-ffi generated rln.h, V3 API) compile
    with -Wall -Wextra -Werror against include/rln_b200.h, link with -lrln_b200 and run.  They are compiled where they lie in
    the reference checkout (never copied); the only shim is an `rln.h` that includes our header.  Without a GPU each one must
    stop at its first step with the library's 'no usable CUDA device' error, not crash."""
    import torch
    (tmp_path / "rln.h").write_text('#include "rln_b200.h"\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(ffi.LIB_PATH)
    exe = str(tmp_path / example)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", str(tmp_path), "-I", os.path.join(root, "include"),
                           "-I", REF_EXAMPLES, os.path.join(REF_EXAMPLES, example + ".c"), "-L", libdir, "-lrln_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    # the examples open ../resources/tree_depth_20/… relative to their working directory
    out = subprocess.run([exe], capture_output=True, text=True, cwd=libdir, timeout=600)
    if torch.cuda.is_available():
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    else:
        assert out.returncode == 1 and "no usable CUDA device" in out.stderr, out.stdout[-2000:] + out.stderr[-2000:]


REF_BUILT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "examples")


@pytest.mark.gpu
@pytest.mark.parametrize("example", ["basic_proof", "multi_message_id", "partial_proof", "recover_secret", "stateless", "type_serialization"])
def test_reference_c_examples_run_on_gpu(example):
    """The reference's own C programs (rln/ffi_c_examples/*.c: tree + proof + verify, multi-message-id, partial proofs, secret
    recovery, stateless mode, type serialisation), compiled UNMODIFIED in this container by oracle/build_ref_examples.py (called
    from __graft_entry__.build()) and shipped as binaries, run to completion against librln_b200.so on the GPU box."""
    exe = os.path.join(REF_BUILT, example)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/examples not built (needs /root/reference at build time)")
    libdir = os.path.dirname(ffi.LIB_PATH)       # the programs open ../resources/tree_depth_20/… relative to their cwd
    env = dict(os.environ, RLN_B200_WINDOW_BITS="8", RLN_B200_WINDOW_BITS_G2="10")
    out = subprocess.run([exe], capture_output=True, text=True, cwd=libdir, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
