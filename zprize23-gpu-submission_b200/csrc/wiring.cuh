// Device-side construction of the sigma polynomials' evaluations from the circuit's wire map (interface of wiring.cu).
#pragma once
#include "common.cuh"
#include "ntt.cuh"
#include "msm.cuh"

namespace zp {

struct WiringScratch {
    DevBuf<uint32_t> keys2, vals2;         // ping-pong buffers of the radix sort
    DevBuf<uint32_t> hist, offs, tile_sum; // per-(digit, tile) counts and their exclusive scan
    DevBuf<uint32_t> first;                // first entry of every variable in the grouped list
};

// Stable LSD radix sort of m (key, value) pairs of u32 by the low `key_bits` bits of the key.  The input arrays are used
// as one side of the ping-pong; *keys_sorted / *vals_sorted point at whichever side holds the result.
void radix_sort_stable_u32(WiringScratch& S, uint32_t* keys, uint32_t* vals, size_t m, int key_bits, uint32_t** keys_sorted,
                           uint32_t** vals_sorted, cudaStream_t st);

// sigma[k][i] = encoding K_w * omega^g of the cell that follows cell (wire k, gate i) in its variable's cycle
// (permutation/mod.rs:101-166).  vars / cells: m device entries in the reference's insertion order, cell = (gate << 2) |
// wire, variable ids < n_vars; both arrays are clobbered (sort workspace).  sigma[k]: N = 2^logn elements each.
void sigma_from_wiring(WiringScratch& S, uint32_t* vars, uint32_t* cells, size_t m, uint32_t n_vars, int logn, const NttTables& T,
                       fr_t* const sigma[4], cudaStream_t st);

}  // namespace zp
