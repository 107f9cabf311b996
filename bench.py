#!/usr/bin/env python3
"""bench.py — RLN proofs/sec on B200 (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the reference arm: CPU restatement on host cores)

A step = one pass of the hot path (witness graph → QAP/NTT → 5 MSMs → assembly → proof bytes) over one
batch of 4 096 synthetic RLN witnesses per GPU (BASELINE.json configs[3]: "batch 4096 RLN proofs,
tree_height=20, 1 B200"); for N GPUs every rank proves its own 4 096 (weak scaling, no data-path
collective; NCCL scatters the witness inputs and gathers the proof bytes outside the kernel path).

  value   whole-job proofs/s with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e     the same through the host C-ABI call rlnb200_prove_batch (host witness bytes in, host proof
          bytes out; H2D/D2H inside the timed region)
  roofline / cpu_baseline as specified in DESIGN.md §measurement
The oracle (oracle/) is used only as the checker of sampled proofs and as the cpu_baseline / reference arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

DEPTH = 20
BATCH = int(os.environ.get("RLN_BENCH_BATCH", "4096"))  # per GPU
METRIC = "rln_proofs_per_sec_batch4096_depth20"
UNIT = "proofs/s"


# The contract is ONE JSON line on stdout.  Libraries loaded later (NCCL prints its version banner to stdout on some boxes)
# must not be able to add lines: fd 1 is pointed at stderr for the whole run and the result line goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------- inputs
def make_witnesses(rln, n, seed, reps=1):
    """member j of a 2^20-leaf tree, message_id = j mod 100, x / (r,s) seeded.
    Returns (reps·n witness records LE, rs bytes, root); record g uses member g mod n with its own x, r, s.
    Uses only the product API + host byte shuffling."""
    import numpy as np
    import zerokit_b200 as z
    from common import fr_stream, fr_bytes, ints, witness_le
    fs = fr_stream(seed)
    limit = 100
    secrets = [next(fs) for _ in range(n)]
    # rate commitments of the n provers (GPU Poseidon through the ABI), random canonical values elsewhere
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 256, size=(1 << DEPTH, 32), dtype=np.uint8)
    raw[:, 31] &= 0x1f
    leaves = bytearray(raw.tobytes())
    for j, s in enumerate(secrets):
        rc = z.poseidon_hash_pair(z.poseidon_hash([s]), limit)
        leaves[32 * j:32 * j + 32] = rc.to_bytes(32, "little")
    rln.set_tree(DEPTH)
    rln.set_leaves_from_bytes(0, bytes(leaves))
    el, bits = rln.get_merkle_proofs(list(range(n)))
    en = z.poseidon_hash_pair(z.hash_to_field_le(b"test-epoch"), z.hash_to_field_le(b"test-rln-identifier"))
    recs, rs = [], []
    paths = [(ints(el[j * DEPTH * 32:(j + 1) * DEPTH * 32]), list(bits[j * DEPTH:(j + 1) * DEPTH])) for j in range(n)]
    for g in range(reps * n):
        j = g % n
        recs.append(witness_le(secrets[j], limit, j % 100, paths[j][0], paths[j][1], next(fs), en))
        rs += [next(fs), next(fs)]
    return b"".join(recs), fr_bytes(rs), rln.get_root()


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
def cpu_prove_sample(ctx, C, inputs, rs, n, threads):
    t = time.perf_counter()
    ctx.prove_batch(inputs[:n * ctx.inputs_size * 32], rs[:64 * n], n, threads)
    return time.perf_counter() - t


def oracle_inputs_from_records(ctx, recs, n):
    """first n witness records → the oracle's input-slot buffers (byte shuffling only)"""
    from common import ints
    rec = 1 + 32 * (5 + DEPTH) + 16 + DEPTH  # rln_witness_to_bytes_le record length (witness.rs:369-415)
    out = []
    for j in range(n):
        b = recs[rec * j:rec * (j + 1)]
        secret, limit, mid = ints(b[1:97])
        pe = ints(b[105:105 + 32 * DEPTH])
        ix = list(b[113 + 32 * DEPTH:113 + 33 * DEPTH])
        x, en = ints(b[113 + 33 * DEPTH:113 + 33 * DEPTH + 64])
        out.append(ctx.inputs_buffer(secret, limit, mid, pe, ix, x, en))
    return b"".join(out)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path.  zerokit is pure Rust and this image has no cargo/rustc,
    so oracle/_ref cannot exist; what is timed is the C++ restatement of the ark-groth16/ark-circom path
    (kind 'port'), one worker thread per proof on all host cores (rln/README.md:324-332)."""
    if rank != 0:
        return
    from oracle import cref_binding as C
    from common import resource, fr_stream, fr_bytes
    C.build()
    threads = C.threads()
    ctx = C.Ctx(resource(DEPTH, "rln_final.arkzkey"), resource(DEPTH, "graph.bin"))
    # same generator as the GPU arm, but paths from the oracle tree would need the 2^20 build; the cost of a
    # proof does not depend on the path values, so a structurally valid witness over a sparse tree is used
    from pyref import poseidon as P
    fs = fr_stream(5)
    n = max(threads * 2, 8)
    inputs, rs = [], []
    tr_el = [P.poseidon([i + 7]) for i in range(DEPTH)]
    for j in range(n):
        inputs.append(ctx.inputs_buffer(next(fs), 100, j % 100, tr_el, [(j >> i) & 1 for i in range(DEPTH)], next(fs), 12345))
        rs += [next(fs), next(fs)]
    inputs, rs = b"".join(inputs), fr_bytes(rs)
    for _ in range(args.warmup):
        cpu_prove_sample(ctx, C, inputs, rs, min(n, threads), threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_prove_sample(ctx, C, inputs, rs, n, threads)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u256-modular",
        "data": "synthetic", "config": {"workload": f"batch {BATCH} RLN proofs, tree_depth=20, bundled zkey (bounded sample of {n} proofs per step)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} proofs per step, one worker thread per proof, C++ restatement of the ark-groth16/ark-circom path"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    dist_on = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if dist_on:
        dist.init_process_group("nccl", device_id=dev)
    import zerokit_b200 as z
    from zerokit_b200 import ffi
    from common import ints
    z.set_device(local_rank)
    t0 = time.time()
    rln = z.RLN.new(DEPTH)
    info = rln.table_info()
    log(f"[rank {rank}] RLN.new: {time.time() - t0:.1f}s, tables G1 c={info['window_bits']} K={info['windows']}{' x2 (GLV)' if info['glv'] else ''}, G2 c={info['window_bits_g2']} K={info['windows_g2']}{' x2 (GLV)' if info['glv'] else ''}, {info['table_bytes'] / 2**30:.1f} GiB")
    n = BATCH
    slots = rln.input_slots()
    # ---- inputs: rank 0 generates world·n distinct witnesses; NCCL scatters contiguous slices (zerokit_b200/sharding.py)
    from zerokit_b200.sharding import scatter_records, gather_records
    rec_len = 1 + 32 * (5 + DEPTH) + 16 + DEPTH
    total = n * world
    full_slots = full_rs = full_recs = None
    if rank == 0:
        t0 = time.time()
        recs_all, rs_all, root = make_witnesses(rln, n, seed=5, reps=world)
        log(f"[rank 0] {total} witnesses + 2^20 tree: {time.time() - t0:.1f}s")
        slot_bytes = b"".join(rln.witness_to_input_slots(recs_all[rec_len * j:rec_len * (j + 1)]) for j in range(total))
        full_slots = torch.frombuffer(bytearray(slot_bytes), dtype=torch.uint8).pin_memory()
        full_rs = torch.frombuffer(bytearray(rs_all), dtype=torch.uint8).pin_memory()
        full_recs = torch.frombuffer(bytearray(recs_all), dtype=torch.uint8).pin_memory()
    if dist_on:
        d_inputs = scatter_records(full_slots, slots * 32, total, dev).contiguous()
        d_rs = scatter_records(full_rs, 64, total, dev).contiguous()
        recs = scatter_records(full_recs, rec_len, total, dev).cpu().numpy().tobytes()   # host records for the e2e leg
        rs = d_rs.cpu().numpy().tobytes()
    else:
        d_inputs = full_slots.to(dev)
        d_rs = full_rs.to(dev)
        recs, rs = recs_all, rs_all
    d_proofs = torch.empty(n * 128, dtype=torch.uint8, device=dev)
    d_values = torch.empty(n * 160, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    rln.reserve(n)

    def step():
        rln.prove_batch_device(d_inputs.data_ptr(), d_rs.data_ptr(), n, d_proofs.data_ptr(), d_values.data_ptr(), 0, stream.cuda_stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    if dist_on:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ffi.lib().rlnb200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc = {}
    torch.cuda.synchronize(dev)
    e0.record(stream)
    for _ in range(args.steps):
        step()
        for k, v in rln.last_stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
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
    value = world * n * args.steps / (ms_max * 1e-3)
    stage = {k: v / args.steps for k, v in stage_acc.items()}

    if args.profile:
        if rank == 0:
            emit({"profile_run": True, "value": value, "stage_ms": stage})
        if dist_on:
            dist.destroy_process_group()
        return
    # ---- gather the 288-byte proof records on rank 0 (NCCL), in global order
    out_proofs = torch.cat([d_proofs.view(n, 128), d_values.view(n, 160)], dim=1).contiguous().view(-1)
    gathered = gather_records(out_proofs, 288, total) if dist_on else out_proofs

    # ---- e2e: host bytes in → host bytes out through the C ABI, every rank on its own slice
    e2e_steps = max(1, min(args.steps, 3))
    rln.prove_batch(recs, n, rs)
    torch.cuda.synchronize(dev)
    if dist_on:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_out = rln.prove_batch(recs, n, rs)
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist_on:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(te.item())

    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return

    # ---- rank 0: correctness of what was timed (oracle = checker only)
    from oracle import cref_binding as C
    from common import resource
    from pyref import groth16 as G
    C.build()
    threads = C.threads()
    ctx = C.Ctx(resource(DEPTH, "rln_final.arkzkey"), resource(DEPTH, "graph.bin"))
    n_chk = 32
    o_inputs = oracle_inputs_from_records(ctx, recs, n_chk)
    want_p, want_pub = ctx.prove_batch(o_inputs, rs[:64 * n_chk], n_chk, threads)
    got = gathered[:288 * n_chk].cpu().numpy().tobytes()
    if dist_on:   # the last rank's slice arrived in order: its first record equals what that rank would print
        assert gathered.numel() == 288 * total
    for j in range(n_chk):
        v = ints(want_p[256 * j:256 * (j + 1)])
        proof = ((v[0], v[1]), ((v[2], v[3]), (v[4], v[5])), (v[6], v[7]))
        y, rt, nul, x, en = ints(want_pub[160 * j:160 * (j + 1)])
        want = G.proof_to_bytes(proof) + b"".join(i.to_bytes(32, "little") for i in (rt, en, x, y, nul))
        assert got[288 * j:288 * (j + 1)] == want, f"proof {j} differs from the oracle"
        assert host_out[290 * j + 1:290 * j + 129] == want[:128], f"host-path proof {j} differs from the oracle"
    ok = rln.verify_batch(host_out, n)
    assert ok == [1] * n, "a proof of the timed batch does not verify"
    t_v = time.perf_counter()
    rln.verify_batch(host_out, n)
    verify_batch_ms = 1e3 * (time.perf_counter() - t_v)
    log(f"[rank 0] checked: first {n_chk} proofs bit-equal to the oracle, all {n} verify")

    # ---- roofline of the dominant kernel: k_msm_accum<Fq> (fixed-base G1 accumulate)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    terms = info["g1_bases"] * n                                     # MSM terms one launch processes
    alg_bytes = 96 * terms                                           # 32 B scalar + 64 B affine base per G1 MSM term
    table_bytes = terms * (32 + 64 * info["adds_per_term"])          # bytes this formulation must touch (scalar + one table entry per window visit)
    k_ms = stage["msm_g1_accum"]
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    mul_rate = z.mul_throughput(2000)
    madds = terms * info["adds_per_term"]
    # wide 32x32->64 MADs of one mixed XYZZ addition as compiled (cuobjdump): 6 products (129) + 2 squarings (100) + the
    # two-term dot product of Y3 (193); IMAD.WIDE occupies the FMA-heavy pipe 4 cycles per warp instruction
    WIDE_PER_ADD = 6 * 129 + 2 * 100 + 193
    sm_clock = 1e6 * float(peaks.get("sm_max_mhz", 1965.0))
    wide_ceiling = 148 * 4 * 8 * sm_clock
    roofline = {
        "bound": "hbm", "kernel": "k_msm_accum<Fq> (fixed-base G1 accumulate)", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src, "kernel_ms": k_ms,
        "algorithmic_bytes_per_launch": alg_bytes, "units_per_launch": terms, "bytes_per_unit": 96,
        "table_formulation_bytes_per_launch": table_bytes, "table_formulation_gbs": table_bytes / (k_ms * 1e-3) / 1e9,
        # the honest second number: this kernel is bound by the integer multiply pipe, not by HBM
        "int_pipe": {"mixed_adds_per_launch": madds, "wide_mads_per_add": WIDE_PER_ADD, "wide_mads_per_s": madds * WIDE_PER_ADD / (k_ms * 1e-3),
                     "wide_mad_ceiling_per_s": wide_ceiling, "frac": madds * WIDE_PER_ADD / (k_ms * 1e-3) / wide_ceiling,
                     "ceiling": "148 SMs x 4 schedulers x 8 lanes/clk (IMAD.WIDE = 4 cycles per warp instruction) x SM clock",
                     "modmul_per_s_measured_peak": mul_rate},
    }
    traffic_file = os.path.join(ROOT, "profiles", "traffic_msm_accum_g1.json")
    if os.path.exists(traffic_file):
        try:
            tj = json.load(open(traffic_file))
            if tj.get("batch") == n and tj.get("window_bits") == info["window_bits"]:
                roofline["traffic"] = tj["dram_bytes_per_launch"]
                roofline["traffic_source"] = tj.get("source")
        except (OSError, ValueError):
            pass

    # ---- CPU baseline on this box's host cores (bounded sample)
    n_cpu = max(4 * threads, 8)
    cpu_in = oracle_inputs_from_records(ctx, recs, min(n_cpu, n))
    n_cpu = min(n_cpu, n)
    cpu_prove_sample(ctx, C, cpu_in, rs, min(threads, n_cpu), threads)
    dt_cpu = cpu_prove_sample(ctx, C, cpu_in, rs, n_cpu, threads)
    cpu_baseline = {"value": n_cpu / dt_cpu, "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": f"{n_cpu} proofs of the same batch, one worker thread per proof on {threads} host threads; C++ restatement of the "
                              f"ark-groth16/ark-circom path (oracle/cref), {dt_cpu:.1f} s"}

    # ---- the other rows of BASELINE.md §3 on the CPU restatement (bounded: a few seconds in total)
    cpu_rows = {}
    try:
        import numpy as np
        t0 = time.perf_counter()
        ctx.prove_batch(cpu_in[:len(cpu_in) // n_cpu], rs[:64], 1, 1)
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

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u256-modular (8x32-bit Montgomery limbs, BN254 Fr/Fq)", "data": "synthetic",
        "config": {"workload": f"batch {n} RLN proofs per GPU, tree_depth=20, bundled zkey (BASELINE.json configs[3])",
                   "global_batch": n * world, "parallelism": f"dp{world} (independent proofs, no data-path collective)",
                   "window_bits": info["window_bits"], "window_bits_g2": info["window_bits_g2"], "table_gib": round(info["table_bytes"] / 2**30, 1),
                   "l2": "working set (tables + 6.5 GB of per-batch matrices) exceeds the 126 MB L2 between iterations"},
        "clocks": clocks, "gpu_launches": int(launches) * world,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * (slots * 32 + 64), "d2h_bytes_per_step": n * 288,
                "steps": e2e_steps, "api": "rlnb200_prove_batch (host witness records → host rln_proof bytes)"},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "stage_ms": stage,
    }
    line["cpu_rows"] = cpu_rows
    # ---- batch verification of the timed batch's proofs (rlnb200_verify_batch: host records in, flags out)
    line["verify_batch"] = {"proofs": n, "ms": verify_batch_ms, "proofs_per_s": n / (verify_batch_ms * 1e-3),
                            "api": "rlnb200_verify_batch (decompression + subgroup checks + 4 Miller loops + final exponentiation per proof)"}
    # ---- single proof through the reference's own entry point (BASELINE.json configs[0]: ffi_generate_rln_proof)
    try:
        wit = z.RLNWitnessInput.from_bytes_le(recs[:rec_len])
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
    except Exception as e:
        line["single_proof"] = {"error": str(e)}
    # ---- two-phase proving (rln/README.md:356-375): partial proofs computed once, finish per message
    try:
        d_pa = torch.empty(n * 320, dtype=torch.uint8, device=dev)
        d_pc = torch.empty(n * 160, dtype=torch.uint8, device=dev)
        d_p2 = torch.empty(n * 128, dtype=torch.uint8, device=dev)
        rln.partial_batch_device(d_inputs.data_ptr(), n, d_pa.data_ptr(), d_pc.data_ptr(), stream.cuda_stream)
        pe0, pe1, pe2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        pe0.record(stream)
        rln.partial_batch_device(d_inputs.data_ptr(), n, d_pa.data_ptr(), d_pc.data_ptr(), stream.cuda_stream)
        pe1.record(stream)
        for _ in range(3):
            rln.finish_batch_device(d_inputs.data_ptr(), d_rs.data_ptr(), d_pa.data_ptr(), n, d_p2.data_ptr(), 0, stream.cuda_stream)
        pe2.record(stream)
        torch.cuda.synchronize(dev)
        assert torch.equal(d_p2, d_proofs), "finish(partial) differs from the full proofs"
        line["two_phase"] = {"partial_proofs_per_s": n / (pe0.elapsed_time(pe1) * 1e-3), "finish_proofs_per_s": 3 * n / (pe1.elapsed_time(pe2) * 1e-3),
                             "note": "finish output bit-equal to the full-proof output of the timed batch (same r, s)"}
    except Exception as e:
        line["two_phase"] = {"error": str(e)}
    if world == 1 and not args.no_micro:
        try:
            line["merkle_microbench"] = merkle_microbench(rln, dev, hbm_peak)
        except Exception as e:
            line["merkle_microbench"] = {"error": str(e)}
    if world == 1 and not args.no_micro:
        try:
            sweep = msm_microbench(z, dev, hbm_peak, sorted(set([16, 18, 20, 22, 24, args.msm_log2])))
            line["msm_g1_microbench"] = next(r for r in sweep if r["log2_n"] == args.msm_log2)
            line["msm_g1_sweep"] = sweep   # BASELINE.json configs[1]: 2^16 … 2^24
        except Exception as e:  # the headline line must still be printed
            line["msm_g1_microbench"] = {"error": str(e)}
    emit(line)
    if dist_on:
        dist.destroy_process_group()


def merkle_microbench(rln, dev, hbm_peak):
    """BASELINE.json configs[2]: Poseidon Merkle tree build over 2^20 leaves resident in HBM + 4 096 membership paths.
    Algorithmic bytes: 64 MiB per build (32 MiB leaves read + 32 MiB nodes written), 660 B per path."""
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
    rln.get_merkle_proofs(idx)
    t_paths = time.perf_counter() - t0
    gbs = (64 << 20) / (ms * 1e-3) / 1e9
    return {"leaves": n, "build_ms": ms, "hashes_per_s": (n - 1) / (ms * 1e-3), "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak,
            "bytes_per_build": 64 << 20, "paths_4096_ms_host_roundtrip": 1e3 * t_paths}


def msm_microbench(z, dev, hbm_peak, sizes):
    """BASELINE.json configs[1]: variable-base G1 MSM, 2^k random scalars for every k in `sizes`; the bases k_i·G are generated
    on the GPU once for the largest size and prefixes of them serve the smaller ones."""
    import torch
    import numpy as np
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
    d_out = torch.empty(64, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev)
    m.gen_bases(d_k.data_ptr(), nmax, d_bases.data_ptr(), st.cuda_stream)
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
        gbs = 96 * n / (ms * 1e-3) / 1e9
        out.append({"log2_n": lg, "ms": ms, "mterms_per_s": n / ms / 1e3, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak,
                    "bytes_per_term": 96})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--msm-log2", type=int, default=22)
    ap.add_argument("--profile", action="store_true", help="timed loop only (for runs under ncu): no e2e, checks, CPU baseline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
