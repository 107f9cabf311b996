// Witness VM translation unit: the kernel body of k_witness_body.cuh with the low-latency Montgomery product (fp.cuh mul_lowlat).
// Measured on a B200 (profiles/r02j_*): cycles per dependent product in a lone warp and the kernel's time with either multiplier.
#define ZK_MUL_LOWLAT 1
#include "k_witness_body.cuh"
