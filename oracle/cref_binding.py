"""ORACLE (test infrastructure): ctypes binding of oracle/_build/liboracle.so."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = ctypes.CDLL(_LIB)
        vp, sz, i32, u8p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p
        L.orc_ctx_new.restype = vp
        L.orc_ctx_new.argtypes = [u8p, sz, u8p, sz]
        L.orc_ctx_free.argtypes = [vp]
        for f in ("orc_ctx_depth", "orc_ctx_inputs_size", "orc_ctx_num_wires", "orc_ctx_domain", "orc_ctx_num_public"):
            getattr(L, f).restype = ctypes.c_uint32
            getattr(L, f).argtypes = [vp]
        L.orc_ctx_input.argtypes = [vp, u8p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
        L.orc_poseidon.argtypes = [u8p, i32, vp]
        L.orc_poseidon_pairs.argtypes = [vp, sz, vp, i32]
        L.orc_poseidon_constants.argtypes = [i32, vp, vp]
        L.orc_merkle_build.argtypes = [ctypes.c_uint32, vp, sz, sz, vp, i32]
        L.orc_witness.argtypes = [vp, vp, vp]
        L.orc_qap_h.argtypes = [vp, vp, vp]
        L.orc_prove_batch.argtypes = [vp, sz, vp, vp, vp, vp, i32]
        L.orc_verify_batch.argtypes = [vp, sz, vp, vp, sz, vp, i32]
        L.orc_msm_g1.argtypes = [vp, vp, sz, vp, i32]
        L.orc_msm_g2.argtypes = [vp, vp, sz, vp, i32]
        L.orc_g1_mul_gen.argtypes = [vp, sz, vp, i32]
        L.orc_fr_dot.argtypes = [vp, vp, sz, vp, i32]
        L.orc_bench_primitive.restype = ctypes.c_double
        L.orc_bench_primitive.argtypes = [i32, i32]
        L.orc_ctx_g1_vec.restype = sz
        L.orc_ctx_g1_vec.argtypes = [vp, i32, vp]
        L.orc_ctx_g2_vec.restype = sz
        L.orc_ctx_g2_vec.argtypes = [vp, vp]
        L.orc_ntt.argtypes = [vp, sz, i32]
        _lib = L
    return _lib


def threads():
    """worker threads for the CPU arm: hardware threads, capped by the container's cgroup CPU quota
    (running more runnable threads than the quota only adds throttling stalls)"""
    n = lib().orc_hardware_threads()
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period) + 0.5)))
    except (OSError, ValueError):
        pass
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    return n


def fr_bytes(vals):
    return b"".join((int(v) % R).to_bytes(32, "little") for v in vals)


def to_ints(buf, n=None):
    b = bytes(buf)
    n = len(b) // 32 if n is None else n
    return [int.from_bytes(b[32 * i:32 * i + 32], "little") for i in range(n)]


def poseidon(vals):
    out = ctypes.create_string_buffer(32)
    lib().orc_poseidon(fr_bytes(vals), len(vals), out)
    return int.from_bytes(out.raw, "little")


def poseidon_constants(t):
    rf_rp = {2: 64, 3: 65, 4: 64, 5: 68, 6: 68, 7: 71, 8: 72, 9: 71}[t]
    ark = ctypes.create_string_buffer(32 * rf_rp * t)
    mds = ctypes.create_string_buffer(32 * t * t)
    lib().orc_poseidon_constants(t, ark, mds)
    return to_ints(ark), to_ints(mds)


def merkle_build(depth, leaves_bytes, start, count, nthreads=1):
    """returns bytes of all 2^(depth+1)-1 nodes (heap order, 32 B LE each)"""
    out = ctypes.create_string_buffer(32 * ((2 << depth) - 1))
    lib().orc_merkle_build(depth, leaves_bytes, start, count, out, nthreads)
    return out.raw


def merkle_proof_from_nodes(nodes, depth, index):
    n = (1 << depth) - 1 + index
    elems, bits = [], []
    while n > 0:
        sib = n + 1 if n & 1 else n - 1
        elems.append(int.from_bytes(nodes[32 * sib:32 * sib + 32], "little"))
        bits.append(0 if n & 1 else 1)
        n = (n - 1) // 2
    return elems, bits


class Ctx:
    def __init__(self, zkey_bytes, graph_bytes):
        self.h = lib().orc_ctx_new(zkey_bytes, len(zkey_bytes), graph_bytes, len(graph_bytes))
        if not self.h:
            raise ValueError("oracle: cannot parse zkey/graph")
        L = lib()
        self.depth = L.orc_ctx_depth(self.h)
        self.inputs_size = L.orc_ctx_inputs_size(self.h)
        self.num_wires = L.orc_ctx_num_wires(self.h)
        self.domain = L.orc_ctx_domain(self.h)
        self.num_public = L.orc_ctx_num_public(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ctx_free(self.h)
            self.h = None

    def input_slot(self, name):
        off, ln = ctypes.c_uint32(), ctypes.c_uint32()
        if not lib().orc_ctx_input(self.h, name.encode(), ctypes.byref(off), ctypes.byref(ln)):
            raise KeyError(name)
        return off.value, ln.value

    def inputs_buffer(self, secret, limit, message_id, path_elements, path_index, x, ext_null, selector_used=None):
        """rln/src/circuit/iden3calc.rs:106-181 + protocol/witness.rs:832-881 → bytes(inputs_size*32).
        Multi message-id circuits: message_id is a list and selector_used a list of bools."""
        buf = [0] * self.inputs_size
        buf[0] = 1
        named = {"identitySecret": [secret], "userMessageLimit": [limit],
                 "messageId": list(message_id) if selector_used is not None else [message_id],
                 "pathElements": list(path_elements), "identityPathIndex": list(path_index),
                 "x": [x], "externalNullifier": [ext_null]}
        if selector_used is not None:
            named["selectorUsed"] = [int(bool(v)) for v in selector_used]
        for k, vals in named.items():
            off, ln = self.input_slot(k)
            assert ln == len(vals), (k, ln, len(vals))
            for i, v in enumerate(vals):
                buf[off + i] = int(v) % R
        return fr_bytes(buf)

    def witness(self, inputs_bytes):
        out = ctypes.create_string_buffer(32 * self.num_wires)
        if not lib().orc_witness(self.h, inputs_bytes, out):
            raise ValueError("graph evaluation failed")
        return out.raw

    def qap_h(self, w_bytes):
        out = ctypes.create_string_buffer(32 * self.domain)
        lib().orc_qap_h(self.h, w_bytes, out)
        return out.raw

    def prove_batch(self, inputs_bytes, rs_bytes, n, nthreads=1):
        """→ (proofs n×256 B [A|B|C affine canonical], publics n×num_public×32 B: the public wires,
        [y,root,nullifier,x,en] for the single circuit)"""
        proofs = ctypes.create_string_buffer(256 * n)
        pub = ctypes.create_string_buffer(32 * self.num_public * n)
        fails = lib().orc_prove_batch(self.h, n, inputs_bytes, rs_bytes, proofs, pub, nthreads)
        if fails:
            raise ValueError(f"{fails} proofs failed")
        return proofs.raw, pub.raw

    def verify_batch(self, proofs, pub, n, npub=None, nthreads=1):
        npub = self.num_public if npub is None else npub
        ok = ctypes.create_string_buffer(n)
        lib().orc_verify_batch(self.h, n, proofs, pub, npub, ok, nthreads)
        return list(ok.raw)

    def g1_vec(self, which):
        n = lib().orc_ctx_g1_vec(self.h, which, None)
        out = ctypes.create_string_buffer(64 * n)
        lib().orc_ctx_g1_vec(self.h, which, out)
        return out.raw

    def g2_vec(self):
        n = lib().orc_ctx_g2_vec(self.h, None)
        out = ctypes.create_string_buffer(128 * n)
        lib().orc_ctx_g2_vec(self.h, out)
        return out.raw


def msm_g1(points_bytes, scalars_bytes, n, nthreads=1):
    out = ctypes.create_string_buffer(64)
    lib().orc_msm_g1(points_bytes, scalars_bytes, n, out, nthreads)
    return out.raw


def msm_g2(points_bytes, scalars_bytes, n, nthreads=1):
    out = ctypes.create_string_buffer(128)
    lib().orc_msm_g2(points_bytes, scalars_bytes, n, out, nthreads)
    return out.raw


def bench_primitive(what, iters):
    """single-thread ns per Fq product (what = 0) or per G1 mixed addition (what = 1) of this port"""
    return lib().orc_bench_primitive(what, iters)


def fr_dot(ks_bytes, ss_bytes, n, nthreads=1):
    """Σ kᵢ·sᵢ mod r as 32 LE bytes: with bases kᵢ·G an MSM must equal g1_mul_gen(fr_dot(k, s))"""
    out = ctypes.create_string_buffer(32)
    lib().orc_fr_dot(ks_bytes, ss_bytes, n, out, nthreads)
    return out.raw


def g1_mul_gen(ks_bytes, n, nthreads=1):
    out = ctypes.create_string_buffer(64 * n)
    lib().orc_g1_mul_gen(ks_bytes, n, out, nthreads)
    return out.raw


def ntt(vals, inverse=False):
    buf = ctypes.create_string_buffer(fr_bytes(vals), 32 * len(vals))
    lib().orc_ntt(buf, len(vals), 1 if inverse else 0)
    return to_ints(buf.raw)
