#!/usr/bin/env python3
"""bench.py — RLN proofs/sec on B200 (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU over NCCL)
  python bench.py --gpus N --inproc ...                    (N>1 in ONE process: rlnb200_multi_prove_batch, a worker thread per GPU)
  python bench.py --impl reference ...                     (the reference arm: CPU restatement on host cores)

A step = one pass of the hot path (witness records → witness graph → QAP/NTT → 5 MSMs → assembly → proof records) over
ONE GLOBAL BATCH of 65 536 synthetic RLN witnesses at tree depth 20 (BASELINE.json configs[4]), whatever N is — strong
scaling.  Every GPU proves its contiguous share in device batches of 4 096 (configs[3], the library's batch size).

  value   whole-job proofs/s with every rank's share of the wire records already resident in HBM; CUDA events on the launching
          stream, max over ranks
  e2e     the same batch through the public entry points with HOST buffers, copies and collectives inside the timed region:
          N = 1: rlnb200_prove_batch (host records → host proof records);
          N > 1: rank-0 host records → H2D → NCCL scatter → rlnb200_prove_records_device on every rank → NCCL gather → D2H
                 to rank-0 host memory (zerokit_b200.sharding.prove_sharded);
          --inproc: rlnb200_multi_prove_batch
  roofline / cpu_baseline as specified in DESIGN.md §5
The oracle (oracle/) is used only as the checker of sampled proofs and as the cpu_baseline / reference arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

DEPTH = 20
GLOBAL_BATCH = int(os.environ.get("RLN_BENCH_GLOBAL_BATCH", "65536"))
DEVICE_BATCH = 4096                                             # the library's device batch (RLN_B200_MAX_BATCH default)
REC_IN = 1 + 32 * (5 + DEPTH) + 16 + DEPTH                      # rln_witness_to_bytes_le record (witness.rs:369-415): 837 B
REC_OUT = 290                                                   # rln_proof_to_bytes_le record (proof.rs:413-428)
METRIC = "rln_proofs_per_sec_batch4096_depth20"
UNIT = "proofs/s"
CPU_SAMPLE = int(os.environ.get("RLN_BENCH_CPU_SAMPLE", "512"))   # proofs of the same batch the CPU leg proves
REF_SAMPLE = int(os.environ.get("RLN_BENCH_REF_SAMPLE", "256"))   # proofs per step of the reference arm


# The contract is ONE JSON line on stdout.  Libraries loaded later (NCCL prints its version banner to stdout on some boxes)
# must not be able to add lines: fd 1 is pointed at stderr for the whole run and the result line goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------- inputs
def make_witnesses(rln, n, seed):
    """member j of a 2^20-leaf tree (leaf j = Poseidon(Poseidon(secret_j), 100), the other leaves seeded
    random), message_id = j mod 100, x and (r, s) seeded, external_nullifier = Poseidon(H("test-epoch"), H("test-rln-identifier"))
    as in rln/benches/partial_proof.rs:22-24.  Returns (n witness records as one uint8 array [n, REC_IN], rs [n, 64], root).
    Uses only the product API (batched GPU Poseidon, the HBM tree) + numpy byte shuffling."""
    import numpy as np
    import zerokit_b200 as z
    limit = 100
    rng = np.random.default_rng(seed)

    def fr_array(count):          # uniform below 2^253 < r: canonical
        a = rng.integers(0, 256, size=(count, 32), dtype=np.uint8)
        a[:, 31] &= 0x1f
        return a
    secrets = fr_array(n)
    idc = np.frombuffer(z.poseidon_hash_batch(secrets.tobytes(), 1, n), dtype=np.uint8).reshape(n, 32)
    lim = np.zeros((n, 32), dtype=np.uint8)
    lim[:, 0] = limit
    rate = np.frombuffer(z.poseidon_hash_batch(np.concatenate([idc, lim], axis=1).tobytes(), 2, n), dtype=np.uint8).reshape(n, 32)
    leaves = np.random.default_rng(3).integers(0, 256, size=(1 << DEPTH, 32), dtype=np.uint8)
    leaves[:, 31] &= 0x1f
    leaves[:n] = rate
    rln.set_tree(DEPTH)
    rln.set_leaves_from_bytes(0, leaves.tobytes())
    el, bits = rln.get_merkle_proofs(list(range(n)))
    el = np.frombuffer(el, dtype=np.uint8).reshape(n, 32 * DEPTH)
    bits = np.frombuffer(bits, dtype=np.uint8).reshape(n, DEPTH)
    en = z.poseidon_hash_pair(z.hash_to_field_le(b"test-epoch"), z.hash_to_field_le(b"test-rln-identifier"))
    rec = np.zeros((n, REC_IN), dtype=np.uint8)
    o = 1
    rec[:, o:o + 32] = secrets; o += 32
    rec[:, o] = limit; o += 32
    rec[:, o] = np.arange(n) % 100; o += 32
    rec[:, o] = DEPTH; o += 8
    rec[:, o:o + 32 * DEPTH] = el; o += 32 * DEPTH
    rec[:, o] = DEPTH; o += 8
    rec[:, o:o + DEPTH] = bits; o += DEPTH
    rec[:, o:o + 32] = fr_array(n); o += 32
    rec[:, o:o + 32] = np.frombuffer(en.to_bytes(32, "little"), dtype=np.uint8); o += 32
    assert o == REC_IN
    rs = np.concatenate([fr_array(n), fr_array(n)], axis=1)
    return rec, rs, rln.get_root()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""

    def __init__(self, index):
        self.index = index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [l.strip().split(",") for l in open(self.tmp.name) if l.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, name in enumerate(names):
                if len(r) > 3 + k and "Active" in r[3 + k] and "Not" not in r[3 + k]:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arm
PORT_NOTE = ("C++ restatement of the ark-groth16 / ark-circom path (oracle/cref: ark-ec window rule Pippenger, radix-2 NTT, ark-ff's no-carry "
             "Montgomery product), one worker thread per proof (rln/README.md:324-332); NOT arkworks itself — no cargo/rustc in this image")


def port_primitives(C):
    """what the port's arithmetic costs on this host, so the ratio can be rescaled against another library's figures"""
    return {"ns_per_fq_mul_1_thread": round(C.bench_primitive(0, 3_000_000), 1), "ns_per_g1_mixed_add_1_thread": round(C.bench_primitive(1, 300_000), 1),
            "note": "measured here on a dependent chain.  Commonly quoted ark-ff 0.5 figures for BN254 Fq (x86-64, asm feature) are ≈ 15–25 ns per product, "
                    "i.e. arkworks is expected to be ≈ 1.5–3× faster per core than this port; those figures could not be re-measured offline"}


def oracle_inputs_from_records(ctx, recs, n):
    """first n witness records (bytes) → the oracle's input-slot buffers (byte shuffling only)"""
    from common import ints
    out = []
    for j in range(n):
        b = recs[REC_IN * j:REC_IN * (j + 1)]
        secret, limit, mid = ints(b[1:97])
        pe = ints(b[105:105 + 32 * DEPTH])
        ix = list(b[113 + 32 * DEPTH:113 + 33 * DEPTH])
        x, en = ints(b[113 + 33 * DEPTH:113 + 33 * DEPTH + 64])
        out.append(ctx.inputs_buffer(secret, limit, mid, pe, ix, x, en))
    return b"".join(out)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path.  zerokit is pure Rust and this image has no cargo/rustc, so oracle/_ref cannot
    exist; what is timed is the C++ restatement (kind 'port') on all host threads the container grants."""
    if rank != 0:
        return
    from oracle import cref_binding as C
    from common import resource, fr_stream, fr_bytes
    C.build()
    threads = C.threads()
    ctx = C.Ctx(resource(DEPTH, "rln_final.arkzkey"), resource(DEPTH, "graph.bin"))
    # same witness generator shape as the GPU arm; the cost of a proof does not depend on the path values, so a structurally valid
    # witness over a sparse tree is used instead of building the 2^20-leaf tree on the CPU
    from pyref import poseidon as P
    fs = fr_stream(5)
    n = max(REF_SAMPLE, threads)
    inputs, rs = [], []
    tr_el = [C.poseidon([i + 7]) for i in range(DEPTH)]
    for j in range(n):
        inputs.append(ctx.inputs_buffer(next(fs), 100, j % 100, tr_el, [(j >> i) & 1 for i in range(DEPTH)], next(fs), 12345))
        rs += [next(fs), next(fs)]
    inputs, rs = b"".join(inputs), fr_bytes(rs)
    isz = ctx.inputs_size * 32
    for _ in range(min(args.warmup, 2)):
        ctx.prove_batch(inputs[:threads * isz], rs[:64 * threads], threads, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.prove_batch(inputs, rs, n, threads)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u256-modular",
        "data": "synthetic",
        "config": {"workload": f"batch {GLOBAL_BATCH} RLN proofs, tree_depth=20, bundled zkey (BASELINE.json configs[4]); each step is a bounded sample of "
                               f"{n} proofs of that workload"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"{n} proofs per step on {threads} host threads; " + PORT_NOTE,
                         "primitives": port_primitives(C)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------------------------- GPU arm
def check_records(C, ctx, recs, rs, got, n_chk, threads, what):
    """first n_chk proof records of `got` bit-equal to the oracle's proofs of the same witness records"""
    from common import ints
    from pyref import groth16 as G
    o_inputs = oracle_inputs_from_records(ctx, recs, n_chk)
    want_p, want_pub = ctx.prove_batch(o_inputs, rs[:64 * n_chk], n_chk, threads)
    for j in range(n_chk):
        v = ints(want_p[256 * j:256 * (j + 1)])
        proof = ((v[0], v[1]), ((v[2], v[3]), (v[4], v[5])), (v[6], v[7]))
        y, rt, nul, x, en = ints(want_pub[160 * j:160 * (j + 1)])
        want = G.rln_proof_to_bytes_le(proof, dict(root=rt, external_nullifier=en, x=x, y=y, nullifier=nul))
        assert got[REC_OUT * j:REC_OUT * (j + 1)] == want, f"{what}: proof {j} differs from the oracle"


def oracle_verify_records(C, ctx, recs_out, idx, threads):
    """the ORACLE's pairing verifier on proof records idx (oracle-side decompression of the 128 compressed bytes)"""
    from common import ints, fr_bytes
    from pyref import groth16 as G
    pts, pubs = [], []
    for j in idx:
        rec = recs_out[REC_OUT * j:REC_OUT * (j + 1)]
        a, b, c = G.proof_from_bytes(rec[1:129])
        pts.append(fr_bytes([a[0], a[1], b[0][0], b[0][1], b[1][0], b[1][1], c[0], c[1]]))
        rt, en, x, y, nul = ints(rec[130:290])
        pubs.append(fr_bytes([y, rt, nul, x, en]))
    return ctx.verify_batch(b"".join(pts), b"".join(pubs), len(idx), 5, threads)


def run_gpu(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    dist_on = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if dist_on:
        dist.init_process_group("nccl", device_id=dev)
    import zerokit_b200 as z
    from zerokit_b200 import ffi
    from zerokit_b200.sharding import gather_records, prove_sharded, rln_prove_records_fn, scatter_records, shard_bounds
    z.set_device(local_rank)
    t0 = time.time()
    rln = z.RLN.new(DEPTH)
    info = rln.table_info()
    log(f"[rank {rank}] RLN.new: {time.time() - t0:.1f}s, tables G1 c={info['window_bits']} K={info['windows']}{' x2 (GLV)' if info['glv'] else ''}, "
        f"G2 c={info['window_bits_g2']} K={info['windows_g2']}{' x2 (GLV)' if info['glv'] else ''}, {info['table_bytes'] / 2**30:.1f} GiB")
    total = GLOBAL_BATCH
    lo, hi = shard_bounds(total, world, rank)
    n = hi - lo
    # ---- inputs: rank 0 generates the global batch (host, pinned); one untimed scatter leaves every rank's share of the wire
    # records resident in HBM for the `value` loop; the e2e loop repeats the scatter from host memory inside its timed region
    h_recs = h_rs = None
    if rank == 0:
        t0 = time.time()
        rec_np, rs_np, root = make_witnesses(rln, total, seed=5)
        h_recs = torch.from_numpy(rec_np.reshape(-1)).pin_memory()
        h_rs = torch.from_numpy(rs_np.reshape(-1)).pin_memory()
        log(f"[rank 0] {total} witness records + 2^20-leaf tree: {time.time() - t0:.1f}s")
    if dist_on:
        d_recs = scatter_records(h_recs, REC_IN, total, dev).contiguous()
        d_rs = scatter_records(h_rs, 64, total, dev).contiguous()
    else:
        d_recs, d_rs = h_recs.to(dev), h_rs.to(dev)
    d_out = torch.empty(n * REC_OUT, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    rln.reserve(min(n, DEVICE_BATCH))

    def step():
        rln.prove_records_device(d_recs.data_ptr(), d_rs.data_ptr(), n, d_out.data_ptr(), stream.cuda_stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    if dist_on:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ffi.lib().rlnb200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc, batches = {}, 0
    torch.cuda.synchronize(dev)
    e0.record(stream)
    for _ in range(args.steps):
        step()
        for k, v in rln.last_stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
        batches += rln.last_stage_batches()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if dist_on:
        dist.barrier()
    launches = ffi.lib().rlnb200_launch_count() - launches0
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist_on:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = total * args.steps / (ms_max * 1e-3)
    stage_per_batch = {k: v / max(batches, 1) for k, v in stage_acc.items()}      # per kernel launch (device batch of ≤ 4 096)
    lt = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if dist_on:
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    launches_all = int(lt.item())

    if args.profile:
        if rank == 0:
            emit({"profile_run": True, "value": value, "stage_ms_per_device_batch": stage_per_batch})
        if dist_on:
            dist.destroy_process_group()
        return
    value_out = d_out.cpu().numpy().tobytes()          # this rank's share, from the timed loop

    # ---- e2e: host records in → host proof records out, copies and collectives inside the timed region
    e2e_steps = max(1, min(args.steps, 10))
    prove_fn = rln_prove_records_fn(rln, REC_OUT)
    h_out = torch.empty(total * REC_OUT, dtype=torch.uint8).pin_memory() if rank == 0 else None

    def e2e_step():
        if dist_on:
            return prove_sharded(prove_fn, h_recs, h_rs, total, REC_IN, REC_OUT, dev, out=h_out)
        rln.prove_batch(h_recs.data_ptr(), total, h_rs.data_ptr(), out=h_out.data_ptr())
        return h_out
    e2e_step()
    torch.cuda.synchronize(dev)
    if dist_on:
        dist.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    f0.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    f1.record(stream)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - w0
    if dist_on:
        dist.barrier()
    # device clock where the work is stream-ordered (N > 1: copies, collectives and kernels all sit on the current stream);
    # rlnb200_prove_batch runs on the library's own stream and blocks the host, so at N = 1 the host clock is the one that sees it
    dt = f0.elapsed_time(f1) * 1e-3 if dist_on else wall
    te = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist_on:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total * e2e_steps / float(te.item())

    # ---- correctness of what was timed (oracle = checker only): EVERY rank checks the head of its own share against the oracle
    from oracle import cref_binding as C
    from common import resource
    C.build()
    threads = max(1, C.threads() // world)
    ctx = C.Ctx(resource(DEPTH, "rln_final.arkzkey"), resource(DEPTH, "graph.bin"))
    n_chk = 32 if world == 1 else 8
    my_recs = d_recs[:REC_IN * n_chk].cpu().numpy().tobytes()
    my_rs = d_rs[:64 * n_chk].cpu().numpy().tobytes()
    check_records(C, ctx, my_recs, my_rs, value_out, n_chk, threads, f"rank {rank} (device-resident loop)")
    okflag = torch.tensor([1], device=dev)
    if dist_on:
        dist.all_reduce(okflag, op=dist.ReduceOp.MIN)      # a rank that failed its assert never gets here: the job dies loudly
    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return
    host_out = h_out.numpy().tobytes()
    if not dist_on:
        assert host_out == value_out, "host-path records differ from the device-resident loop's"
    threads = C.threads()
    # rank 0: the gathered e2e output, head of EVERY rank's slice bit-equal to the oracle; a sample of every slice under the ORACLE verifier
    all_recs, all_rs = h_recs.numpy().tobytes(), h_rs.numpy().tobytes()
    sample = []
    for r in range(world):
        rlo, rhi = shard_bounds(total, world, r)
        k = min(8, rhi - rlo)
        check_records(C, ctx, all_recs[REC_IN * rlo:REC_IN * (rlo + k)], all_rs[64 * rlo:64 * (rlo + k)], host_out[REC_OUT * rlo:REC_OUT * (rlo + k)],
                      k, threads, f"e2e output, slice of rank {r}")
        step_ = max(1, (rhi - rlo) // max(1, 256 // world))
        sample += list(range(rlo, rhi, step_)) + [rhi - 1]
    assert oracle_verify_records(C, ctx, host_out, sample, threads) == [1] * len(sample), "a sampled proof fails under the oracle verifier"
    t_v = time.perf_counter()
    ok = rln.verify_batch(host_out, total)
    verify_batch_ms = 1e3 * (time.perf_counter() - t_v)
    assert ok == [1] * total, "a proof of the timed batch does not verify"
    log(f"[rank 0] checked: head of every rank's slice bit-equal to the oracle, {len(sample)} sampled proofs under the oracle verifier, all {total} under the GPU verifier")

    # ---- roofline of the dominant kernel: k_msm_accum<Fq> (fixed-base G1 accumulate), per launch = per device batch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    per_launch = min(n, DEVICE_BATCH)
    terms = info["g1_bases"] * per_launch                            # MSM terms one launch processes
    alg_bytes = 96 * terms                                           # 32 B scalar + 64 B affine base per G1 MSM term
    table_bytes = terms * (32 + 64 * info["adds_per_term"])          # bytes this formulation must touch (scalar + one table entry per window visit)
    k_ms = stage_per_batch["msm_g1_accum"]
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    mul_rate = z.mul_throughput(2000)
    madds = terms * info["adds_per_term"]
    # wide 32x32->64 MADs of one mixed XYZZ addition as compiled (cuobjdump): 6 products (129) + 2 squarings (100) + the
    # two-term dot product of Y3 (193); IMAD.WIDE occupies the FMA-heavy pipe 4 cycles per warp instruction
    WIDE_PER_ADD = 6 * 129 + 2 * 100 + 193
    sm_clock = 1e6 * float(peaks.get("sm_max_mhz", 1965.0))
    wide_ceiling = 148 * 4 * 8 * sm_clock
    int_frac = madds * WIDE_PER_ADD / (k_ms * 1e-3) / wide_ceiling
    roofline = {
        "bound": "hbm", "kernel": "k_msm_accum<Fq> (fixed-base G1 accumulate)", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src, "kernel_ms": k_ms,
        "algorithmic_bytes_per_launch": alg_bytes, "units_per_launch": terms, "bytes_per_unit": 96,
        "table_formulation_bytes_per_launch": table_bytes, "table_formulation_gbs": table_bytes / (k_ms * 1e-3) / 1e9,
        # the honest second numbers: this kernel is bound by the integer multiply pipe, not by HBM
        "int_pipe_frac": int_frac, "sm__pipe_fmaheavy_cycles_active_pct": None,
        "int_pipe": {"mixed_adds_per_launch": madds, "wide_mads_per_add": WIDE_PER_ADD, "wide_mads_per_s": madds * WIDE_PER_ADD / (k_ms * 1e-3),
                     "wide_mad_ceiling_per_s": wide_ceiling, "frac": int_frac,
                     "ceiling": "148 SMs x 4 schedulers x 8 lanes/clk (IMAD.WIDE = 4 cycles per warp instruction) x SM clock",
                     "modmul_per_s_measured_peak": mul_rate},
    }
    traffic_file = os.path.join(ROOT, "profiles", "traffic_msm_accum_g1.json")
    if os.path.exists(traffic_file):
        try:
            tj = json.load(open(traffic_file))
            if tj.get("batch") == per_launch and tj.get("window_bits") == info["window_bits"]:
                roofline["traffic"] = tj["dram_bytes_per_launch"]
                roofline["traffic_source"] = tj.get("source")
                roofline["sm__pipe_fmaheavy_cycles_active_pct"] = tj.get("sm__pipe_fmaheavy_cycles_active_pct")
        except (OSError, ValueError):
            pass

    # ---- CPU baseline on this box's host cores: a bounded sample of the same batch
    n_cpu = min(max(CPU_SAMPLE, threads), total)
    cpu_in = oracle_inputs_from_records(ctx, all_recs, n_cpu)
    ctx.prove_batch(cpu_in[:threads * ctx.inputs_size * 32], all_rs[:64 * threads], min(threads, n_cpu), threads)
    t0 = time.perf_counter()
    ctx.prove_batch(cpu_in, all_rs[:64 * n_cpu], n_cpu, threads)
    dt_cpu = time.perf_counter() - t0
    cpu_baseline = {"value": n_cpu / dt_cpu, "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": f"the first {n_cpu} proofs of the same batch on {threads} host threads, {dt_cpu:.1f} s; " + PORT_NOTE,
                    "primitives": port_primitives(C),
                    "ratio_note": f"value / cpu_baseline.value holds for {threads} host threads and this port; it halves on a host with twice the threads"}

    # ---- the other rows of BASELINE.md §3 on the CPU restatement (bounded: a few seconds in total)
    cpu_rows = {}
    try:
        t0 = time.perf_counter()
        ctx.prove_batch(cpu_in[:ctx.inputs_size * 32], all_rs[:64], 1, 1)
        cpu_rows["single_proof_ms_1_thread"] = 1e3 * (time.perf_counter() - t0)
        rng = np.random.default_rng(1)
        ks = rng.integers(0, 256, size=(4096, 32), dtype=np.uint8)
        ks[:, 31] &= 0x1f
        tile = C.g1_mul_gen(ks.tobytes(), 4096, threads)          # 4 096 distinct points, tiled: Pippenger's cost does not depend on the values
        rows = []
        for lg in (16, 18, 20):
            nn = 1 << lg
            sc = rng.integers(0, 256, size=(nn, 32), dtype=np.uint8)
            sc[:, 31] &= 0x1f
            t0 = time.perf_counter()
            C.msm_g1(tile * (nn // 4096), sc.tobytes(), nn, threads)
            rows.append({"log2_n": lg, "ms": 1e3 * (time.perf_counter() - t0)})
        cpu_rows["msm_g1"] = rows
        lv = rng.integers(0, 256, size=(1 << DEPTH, 32), dtype=np.uint8)
        lv[:, 31] &= 0x1f
        t0 = time.perf_counter()
        C.merkle_build(DEPTH, lv.tobytes(), 0, 1 << DEPTH, threads)
        cpu_rows["merkle_build_2pow20_ms"] = 1e3 * (time.perf_counter() - t0)
        cpu_rows["threads"] = threads
        cpu_rows["what"] = "oracle/cref (C++ restatement of the ark path): ark-ec window rule Pippenger, level-parallel FullMerkleTree build"
    except Exception as e:   # noqa: BLE001
        cpu_rows["error"] = str(e)

    par = f"dp{world}: contiguous shards of independent proofs, no collective inside the computation"
    if dist_on:
        par += "; NCCL scatter of witness records / gather of proof records (inside the e2e timed region)"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u256-modular (8x32-bit Montgomery limbs, BN254 Fr/Fq)", "data": "synthetic",
        "config": {"workload": f"batch {total} RLN proofs per step over {world} GPU(s) (BASELINE.json configs[4]), proved in device batches of "
                               f"{min(n, DEVICE_BATCH)} (configs[3]), tree_depth=20, bundled zkey",
                   "global_batch": total, "per_gpu": n, "device_batch": min(n, DEVICE_BATCH), "parallelism": par,
                   "window_bits": info["window_bits"], "window_bits_g2": info["window_bits_g2"], "table_gib": round(info["table_bytes"] / 2**30, 1),
                   "l2": "working set (tables + 6.5 GB of per-batch matrices) exceeds the 126 MB L2 between iterations"},
        "clocks": clocks, "gpu_launches": launches_all,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": total * (REC_IN + 64), "d2h_bytes_per_step": total * REC_OUT,
                "steps": e2e_steps, "clock": "CUDA events on the stream that carries copies, collectives and kernels, max over ranks" if dist_on else "host clock around the blocking C-ABI call",
                "api": ("zerokit_b200.sharding.prove_sharded: rank-0 pinned host records → H2D → NCCL scatter → rlnb200_prove_records_device per rank → NCCL gather → D2H to rank-0 host"
                        if dist_on else "rlnb200_prove_batch (host witness records → host rln_proof records; record parsing and formatting on the GPU)")},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "stage_ms_per_device_batch": stage_per_batch,
    }
    line["cpu_rows"] = cpu_rows
    # ---- batch verification of the timed batch's proofs (rlnb200_verify_batch: host records in, flags out)
    line["verify_batch"] = {"proofs": total, "ms": verify_batch_ms, "proofs_per_s": total / (verify_batch_ms * 1e-3),
                            "api": "rlnb200_verify_batch (decompression + G2 membership + one merged Miller loop + final exponentiation per proof; above 4 096 proofs one thread each, up to 4 096 one CTA each on the lane-parallel program)"}
    try:   # the same check on 1 … 4 096 proofs: the lane-parallel verifier (k_verify_vm) that single calls and small batches take
        small = {}
        for nv in (1, 32, 296, 4096):
            if nv > total:
                continue
            buf = host_out[:REC_OUT * nv]
            assert rln.verify_batch(buf, nv) == [1] * nv
            t_v = time.perf_counter()
            for _ in range(5):
                rln.verify_batch(buf, nv)
            ms = 1e3 * (time.perf_counter() - t_v) / 5
            small[str(nv)] = {"ms": ms, "proofs_per_s": nv / (ms * 1e-3)}
        line["verify_batch"]["lane_parallel"] = dict(small, program=rln.verify_vm_info())
    except Exception as e:
        line["verify_batch"]["lane_parallel"] = {"error": str(e)}
    # ---- single proof / single verification through the reference's own entry points (BASELINE.json configs[0])
    try:
        wit = z.RLNWitnessInput.from_bytes_le(all_recs[:REC_IN])
        rln.generate_rln_proof(wit)
        t0 = time.perf_counter()
        for _ in range(5):
            p1 = rln.generate_rln_proof(wit)
        lat = (time.perf_counter() - t0) / 5
        assert rln.verify_with_roots(p1, p1.values.x, []) is True
        t0 = time.perf_counter()
        for _ in range(5):
            rln.verify_with_roots(p1, p1.values.x, [])
        line["single_proof"] = {"generate_ms": 1e3 * lat, "verify_ms": 1e3 * (time.perf_counter() - t0) / 5,
                                "api": "ffi_generate_rln_proof / ffi_verify_with_roots, host in, host out"}
        line["single_proof"]["concurrent_callers"] = concurrent_callers(z, rln, all_recs)
    except Exception as e:
        line["single_proof"] = {"error": str(e)}
    if world == 1 and not args.no_micro:
        # ---- two-phase proving (rln/README.md:356-375) on one device batch: partial proofs computed once, finish per message
        try:
            nb = DEVICE_BATCH
            slots = rln.input_slots()
            d_in = torch.frombuffer(bytearray(b"".join(rln.witness_to_input_slots(all_recs[REC_IN * j:REC_IN * (j + 1)]) for j in range(nb))), dtype=torch.uint8).to(dev)
            d_pa = torch.empty(nb * 320, dtype=torch.uint8, device=dev)
            d_pc = torch.empty(nb * 160, dtype=torch.uint8, device=dev)
            d_p1 = torch.empty(nb * 128, dtype=torch.uint8, device=dev)
            d_p2 = torch.empty(nb * 128, dtype=torch.uint8, device=dev)
            rln.prove_batch_device(d_in.data_ptr(), d_rs.data_ptr(), nb, d_p1.data_ptr(), 0, 0, stream.cuda_stream)
            rln.partial_batch_device(d_in.data_ptr(), nb, d_pa.data_ptr(), d_pc.data_ptr(), stream.cuda_stream)
            pe0, pe1, pe2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            pe0.record(stream)
            rln.partial_batch_device(d_in.data_ptr(), nb, d_pa.data_ptr(), d_pc.data_ptr(), stream.cuda_stream)
            pe1.record(stream)
            for _ in range(3):
                rln.finish_batch_device(d_in.data_ptr(), d_rs.data_ptr(), d_pa.data_ptr(), nb, d_p2.data_ptr(), 0, stream.cuda_stream)
            pe2.record(stream)
            torch.cuda.synchronize(dev)
            assert torch.equal(d_p2, d_p1), "finish(partial) differs from the full proofs"
            assert d_p1.cpu().numpy().tobytes()[:128] == host_out[1:129]
            line["two_phase"] = {"batch": nb, "partial_proofs_per_s": nb / (pe0.elapsed_time(pe1) * 1e-3), "finish_proofs_per_s": 3 * nb / (pe1.elapsed_time(pe2) * 1e-3),
                                 "note": "finish output bit-equal to the full-proof output of the same device batch (same r, s)"}
            assert slots * 32 * nb == d_in.numel()
        except Exception as e:
            line["two_phase"] = {"error": str(e)}
        try:
            line["merkle_microbench"] = merkle_microbench(rln, dev, hbm_peak, C)
        except Exception as e:
            line["merkle_microbench"] = {"error": str(e)}
        try:
            sweep = msm_microbench(z, dev, hbm_peak, sorted(set([16, 18, 20, 22, 24, args.msm_log2])), C)
            line["msm_g1_microbench"] = next(r for r in sweep if r["log2_n"] == args.msm_log2)
            line["msm_g1_sweep"] = sweep   # BASELINE.json configs[1]: 2^16 … 2^24, every result checked
        except Exception as e:  # the headline line must still be printed
            line["msm_g1_microbench"] = {"error": str(e)}
    emit(line)
    if dist_on:
        dist.destroy_process_group()


def concurrent_callers(z, rln, all_recs, threads=16, rounds=4):
    """many host threads on ONE handle, each making single-item calls (rln/README.md:324-332: the reference's scaling model)"""
    import threading
    wits = [z.RLNWitnessInput.from_bytes_le(all_recs[REC_IN * j:REC_IN * (j + 1)]) for j in range(threads)]
    errs = []

    def work(t):
        try:
            for _ in range(rounds):
                p = rln.generate_rln_proof(wits[t])
                assert rln.verify_with_roots(p, p.values.x, []) is True
        except Exception as e:   # noqa: BLE001
            errs.append(str(e))
    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    if errs:
        return {"error": errs[0]}
    return {"threads": threads, "proofs_per_s": threads * rounds / dt, "verifications_per_s": threads * rounds / dt,
            "note": "each thread: ffi_generate_rln_proof then ffi_verify_with_roots, one handle; concurrent single-item calls are coalesced into device batches"}


def run_inproc(args):
    """N GPUs in ONE process: rlnb200_multi_prove_batch (replica + worker thread per device), host records in, host records out"""
    import torch
    import zerokit_b200 as z
    from zerokit_b200 import ffi
    ndev = args.gpus
    t0 = time.time()
    multi = z.RLNMulti(DEPTH, list(range(ndev)))
    log(f"[inproc] {ndev} replicas: {time.time() - t0:.1f}s")
    total = GLOBAL_BATCH
    r0 = multi.replica(0)
    rec_np, rs_np, root = make_witnesses(r0, total, seed=5)
    multi.reserve(DEVICE_BATCH)
    h_recs = torch.from_numpy(rec_np.reshape(-1)).pin_memory()
    h_rs = torch.from_numpy(rs_np.reshape(-1)).pin_memory()
    h_out = torch.empty(total * REC_OUT, dtype=torch.uint8).pin_memory()
    sampler = ClockSampler(0)
    launches0 = ffi.lib().rlnb200_launch_count()
    for _ in range(args.warmup):
        multi.prove_batch(h_recs.data_ptr(), total, h_rs.data_ptr(), out=h_out.data_ptr())
    sampler.start()
    t0 = time.perf_counter()
    shard = []
    for _ in range(args.steps):
        multi.prove_batch(h_recs.data_ptr(), total, h_rs.data_ptr(), out=h_out.data_ptr())
        shard.append(multi.last_shard_ms())
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = ffi.lib().rlnb200_launch_count() - launches0
    v = total * args.steps / dt
    from oracle import cref_binding as C
    from common import resource
    C.build()
    ctx = C.Ctx(resource(DEPTH, "rln_final.arkzkey"), resource(DEPTH, "graph.bin"))
    host_out, all_recs, all_rs = h_out.numpy().tobytes(), h_recs.numpy().tobytes(), h_rs.numpy().tobytes()
    from zerokit_b200.sharding import shard_bounds
    for r in range(ndev):
        rlo, rhi = shard_bounds(total, ndev, r)
        check_records(C, ctx, all_recs[REC_IN * rlo:REC_IN * (rlo + 8)], all_rs[64 * rlo:64 * (rlo + 8)], host_out[REC_OUT * rlo:REC_OUT * (rlo + 8)], 8, C.threads(),
                      f"slice of device {r}")
    assert multi.verify_batch(host_out, total) == [1] * total
    emit({"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": ndev, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u256-modular (8x32-bit Montgomery limbs, BN254 Fr/Fq)", "data": "synthetic",
          "config": {"workload": f"batch {total} RLN proofs per step over {ndev} GPU(s) in ONE process (rlnb200_multi_prove_batch), tree_depth=20", "global_batch": total,
                     "parallelism": f"in-process dp{ndev}: a replica and a worker thread per device, contiguous shards, no collective"},
          "clocks": clocks, "gpu_launches": int(launches),
          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": total * (REC_IN + 64), "d2h_bytes_per_step": total * REC_OUT, "steps": args.steps,
                  "api": "rlnb200_multi_prove_batch (host records → host records); host clock around the blocking call"},
          "shard_ms_last_step": shard[-1] if shard else None,
          "note": "value == e2e here: this mode only has the host-buffer entry point; heads of every device's slice bit-equal to the oracle, all proofs verify"})


def merkle_microbench(rln, dev, hbm_peak, C):
    """BASELINE.json configs[2]: Poseidon Merkle tree build over 2^20 leaves resident in HBM + 4 096 membership paths, root and a
    path checked against the oracle.  Algorithmic bytes: 64 MiB per build (32 MiB leaves read + 32 MiB nodes written), 660 B per path."""
    import torch
    import numpy as np
    n = 1 << DEPTH
    rng = np.random.default_rng(7)
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    raw[:, 31] &= 0x1f
    d_leaves = torch.from_numpy(raw).to(dev)
    st = torch.cuda.current_stream(dev)
    rln.set_tree(DEPTH)
    rln.set_leaves_from_device(0, d_leaves.data_ptr(), n, st.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    torch.cuda.synchronize(dev)
    e0.record(st)
    for _ in range(reps):
        rln.set_leaves_from_device(0, d_leaves.data_ptr(), n, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    idx = [int(x) for x in rng.integers(0, n, size=4096)]
    t0 = time.perf_counter()
    el, bits = rln.get_merkle_proofs(idx)
    t_paths = time.perf_counter() - t0
    nodes = C.merkle_build(DEPTH, raw.tobytes(), 0, n, C.threads())
    assert rln.get_root() == int.from_bytes(nodes[:32], "little"), "2^20-leaf root differs from the oracle"
    from common import ints
    for k in (0, 1000, 4095):
        e, b = C.merkle_proof_from_nodes(nodes, DEPTH, idx[k])
        assert ints(el[k * 640:(k + 1) * 640]) == e and list(bits[k * 20:(k + 1) * 20]) == b
    gbs = (64 << 20) / (ms * 1e-3) / 1e9
    return {"leaves": n, "build_ms": ms, "hashes_per_s": (n - 1) / (ms * 1e-3), "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak,
            "bytes_per_build": 64 << 20, "paths_4096_ms_host_roundtrip": 1e3 * t_paths, "checked": "root and sampled paths == oracle"}


def msm_microbench(z, dev, hbm_peak, sizes, C):
    """BASELINE.json configs[1]: variable-base G1 MSM, 2^k random scalars for every k in `sizes`; the bases k_i·G are generated on
    the GPU once for the largest size and prefixes of them serve the smaller ones.  EVERY timed result is checked: the bases are
    k_i·G, so the MSM must equal (Σ k_i·s_i mod r)·G (dot product and scalar multiplication by the oracle); up to 2^18 also against
    the oracle's own Pippenger; above, additionally Σ(whole) == Σ(first half) + Σ(second half)."""
    import torch
    import numpy as np
    from common import fr_bytes
    nmax = 1 << max(sizes)
    m = z.G1Msm(nmax)
    rng = np.random.default_rng(1)
    ks = rng.integers(0, 256, size=(nmax, 32), dtype=np.uint8)
    ks[:, 31] &= 0x1f
    sc = rng.integers(0, 256, size=(nmax, 32), dtype=np.uint8)
    sc[:, 31] &= 0x1f
    d_k = torch.from_numpy(ks).to(dev)
    d_s = torch.from_numpy(sc).to(dev)
    d_bases = torch.empty(nmax * 64, dtype=torch.uint8, device=dev)
    d_out = torch.empty(3 * 64, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev)
    m.gen_bases(d_k.data_ptr(), nmax, d_bases.data_ptr(), st.cuda_stream)
    th = C.threads()
    out = []
    for lg in sizes:
        n = 1 << lg
        for _ in range(2):
            m.msm_device(d_bases.data_ptr(), d_s.data_ptr(), n, d_out.data_ptr(), st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if lg <= 22 else 3
        torch.cuda.synchronize(dev)
        e0.record(st)
        for _ in range(reps):
            m.msm_device(d_bases.data_ptr(), d_s.data_ptr(), n, d_out.data_ptr(), st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
        res = d_out[:64].cpu().numpy().tobytes()
        checks = ["(sum k_i s_i) G"]
        assert C.g1_mul_gen(C.fr_dot(ks[:n].tobytes(), sc[:n].tobytes(), n, th), 1) == res, f"MSM 2^{lg}: result != (Σ k_i s_i)·G"
        if lg <= 18:
            bases = C.g1_mul_gen(ks[:n].tobytes(), n, th)
            assert C.msm_g1(bases, sc[:n].tobytes(), n, th) == res, f"MSM 2^{lg}: result differs from the oracle's Pippenger"
            checks.append("oracle Pippenger")
        else:
            h = n // 2
            m.msm_device(d_bases.data_ptr(), d_s.data_ptr(), h, d_out.data_ptr() + 64, st.cuda_stream)
            m.msm_device(d_bases.data_ptr() + 64 * h, d_s.data_ptr() + 32 * h, h, d_out.data_ptr() + 128, st.cuda_stream)
            torch.cuda.synchronize(dev)
            o = d_out.cpu().numpy().tobytes()
            assert C.msm_g1(o[64:192], fr_bytes([1, 1]), 2) == res, f"MSM 2^{lg}: whole != sum of halves"
            checks.append("halves")
        gbs = 96 * n / (ms * 1e-3) / 1e9
        out.append({"log2_n": lg, "ms": ms, "mterms_per_s": n / ms / 1e3, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak,
                    "bytes_per_term": 96, "checked": checks})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--msm-log2", type=int, default=22)
    ap.add_argument("--inproc", action="store_true", help="N GPUs in one process through rlnb200_multi_prove_batch (not under torchrun)")
    ap.add_argument("--profile", action="store_true", help="timed loop only (for runs under ncu): no e2e, checks, CPU baseline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.inproc or (args.gpus > 1 and world == 1):
        run_inproc(args)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
