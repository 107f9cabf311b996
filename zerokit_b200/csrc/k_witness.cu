// Witness VM translation unit (the kernel body is k_witness_body.cuh).  Round 2 compiled it with a "low-latency" Montgomery product
// (separated reduction on 4×4-limb blocks: 164 wide MADs in short chains instead of 129 in two long carry chains) on the theory that
// a lone warp pays dependency depth, not MAD count.  Measured on a B200 (profiles/r02j_latency_probe_lowlat_rejected.txt): cycles
// per dependent product in a lone warp — mul_ptx 865, portable CIOS 1 131, the block form 1 490; witness kernel 6.85 → 9.17 ms.
// The interleaved-carry form is also the lowest-latency one; the block form was deleted.
#include "k_witness_body.cuh"
