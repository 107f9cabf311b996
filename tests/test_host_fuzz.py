"""CPU: robustness of the host-side parsers that take untrusted bytes across the C ABI.

* tests/host_fuzz/fuzz_parsers.cpp — the arkzkey / witnesscalc-graph parsers and the VM scheduler of zerokit_b200/csrc/host_util.hpp
  under -fsanitize=address,undefined: hand-made hostile files (lengths that wrap a 64-bit offset, input signals outside the inputs
  buffer, indices ≥ 2^32) must be rejected, random mutations must parse into in-range structures or throw.
* tests/host_fuzz/fuzz_records.py — every byte-record parser of the shared library (ffi_bytes_{le,be,mixed}_to_*), through ctypes in a
  subprocess so a crash is a test failure rather than the end of the run.

The reference gets the same guarantees from safe Rust (rln/src/utils.rs readers return Err on short input; serde/prost reject
malformed files); here they have to be tested."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES = os.path.join(ROOT, "zerokit_b200", "resources", "tree_depth_20")


@pytest.fixture(scope="module")
def harness():
    out = os.path.join(ROOT, "tests", "host_fuzz", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "fuzz_parsers")
    subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "zerokit_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_fuzz", "fuzz_parsers.cpp"), "-o", exe])
    return exe


@pytest.mark.parametrize("sub,iters", [("", 240), ("multi_message_id/max_out_4", 120)])
def test_file_parsers_sanitized(harness, sub, iters):
    d = os.path.join(RES, sub)
    r = subprocess.run([harness, os.path.join(d, "rln_final.arkzkey"), os.path.join(d, "graph.bin"), str(iters), "11"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "crafted graphs: 10 of 10 rejected" in r.stdout and "fuzz done" in r.stdout


def test_record_parsers_do_not_crash():
    env = dict(os.environ, N="1500", SEED="21")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_fuzz", "fuzz_records.py")], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "fuzz ok" in r.stdout


def test_request_coalescer_thread_sanitizer():
    """zerokit_b200/csrc/coalesce.hpp (concurrent single-item calls on one handle run as batches) under -fsanitize=thread: every
    request answered once with its own result, batch limit respected, the batch function never re-entered, batches really shared"""
    out = os.path.join(ROOT, "tests", "host_fuzz", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "coalesce_test")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", "-I", os.path.join(ROOT, "zerokit_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_fuzz", "coalesce_test.cpp"), "-o", exe])
    for args in (("16", "40", "8"), ("48", "20", "4096"), ("3", "50", "2"), ("1", "20", "8")):
        r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "coalesce ok" in r.stdout and "ThreadSanitizer" not in r.stderr, r.stdout + r.stderr[-3000:]
