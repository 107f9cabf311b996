"""ORACLE — test infrastructure only.

CPU restatements of the reference's RLN Groth16 proving path:
  oracle.pyref   Python-integer version, pinned to the reference's golden vectors (tests/golden)
  oracle.cref    C++ version (oracle/cref → oracle/_build/liboracle.so), validated against pyref

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (zerokit_b200) never does.

Parity status: PINNED — the Python restatement reproduces every Poseidon constant and hash KAT, the
depth-20 tree root/path KAT and accepts the reference's hard-coded snarkjs proof (see
tests/test_oracle_goldens.py).  There is no oracle/_ref: the reference is pure Rust and this
image has no cargo/rustc, so it cannot be compiled here.
"""
