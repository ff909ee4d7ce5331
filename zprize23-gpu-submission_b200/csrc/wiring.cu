// Permutation-argument preprocessing on the device: sigma_k(omega^i) from the circuit's wire map.
//
// Reference: `Permutation::compute_sigma_permutations` + `compute_permutation_lagrange`
// ("Prize 1B/plonk-core/src/permutation/mod.rs":101-166): every variable keeps the list of wire cells it was
// assigned to, in insertion order; each cell maps to the NEXT cell of its variable's list (the last one wraps to the
// first), cells that no variable owns map to themselves; the cell (wire k, gate i) is encoded as K_k * omega^i with
// K = (1, 7, 13, 17) (permutation/constants.rs:12-22).  The reference walks a HashMap<Variable, Vec<WireData>> on one CPU
// thread (3.2 M variables / 12.6 M cells at HEIGHT=15).
//
// Here the flattened map — m entries (variable, cell), cell = (gate << 2) | wire, in insertion order — is grouped by
// variable with a STABLE least-significant-digit radix sort (8-bit digits, only as many passes as the variable ids need),
// which keeps the insertion order inside every group; one more pass links each entry to its successor and writes the
// encoded field element.  Everything is u32 index work bound by HBM bandwidth: 3 passes x (8 B read + 8 B scattered
// write) per entry.
#include "wiring.cuh"

namespace zp {

static const int RS_THREADS = 256;             // 8 warps
static const int RS_ROUNDS = 8;                // elements per thread
static const int RS_TILE = RS_THREADS * RS_ROUNDS;

// lanes of the warp holding the same 8-bit digit (ballot per bit: runs on every architecture and under the emulator)
ZP_D uint32_t same_digit_mask(uint32_t d, bool valid) {
    uint32_t m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; b++) {
        uint32_t bal = __ballot_sync(0xffffffffu, valid && ((d >> b) & 1u));
        m &= ((d >> b) & 1u) ? bal : ~bal;
    }
    return valid ? m : 0u;
}

// hist[digit * ntiles + tile] = number of keys of the tile with that digit
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint32_t* __restrict__ keys, size_t m, int shift, size_t ntiles,
                                                             uint32_t* __restrict__ hist) {
    __shared__ uint32_t cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; r++) {
        size_t i = base + (size_t)r * RS_THREADS + threadIdx.x;
        if (i < m) atomicAdd(&cnt[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = cnt[threadIdx.x];
}

// Stable scatter of one tile.  Element order inside the tile is (warp, round, lane) = its position in the array: warp w
// owns the 256 consecutive elements [w * 256, (w + 1) * 256) and walks them 32 at a time, so "earlier in the array" is
// "earlier (warp, round, lane)" and ranks assigned in that order keep equal keys in input order.
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                                size_t m, int shift, size_t ntiles,
                                                                const uint32_t* __restrict__ offs, uint32_t* __restrict__ keys_out,
                                                                uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t cnt[RS_THREADS / 32][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (RS_THREADS / 32) * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE + (size_t)w * (32 * RS_ROUNDS);
    uint32_t k[RS_ROUNDS], v[RS_ROUNDS], rank[RS_ROUNDS];
    for (int r = 0; r < RS_ROUNDS; r++) {
        const size_t i = base + (size_t)r * 32 + lane;
        const bool valid = i < m;
        k[r] = valid ? keys[i] : 0;
        v[r] = valid ? vals[i] : 0;
        const uint32_t d = (k[r] >> shift) & 255u;
        const uint32_t peers = same_digit_mask(d, valid);
        const int leader = peers ? __ffs(peers) - 1 : 0;
        uint32_t start = 0;
        if (valid && lane == leader) {
            start = cnt[w][d];
            cnt[w][d] = start + __popc(peers);
        }
        start = __shfl_sync(0xffffffffu, start, leader);
        rank[r] = start + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {   // digit = threadIdx.x: turn the per-warp counts into start positions (global offset of (digit, tile) first)
        const uint32_t d = threadIdx.x;
        uint32_t run = offs[(size_t)d * ntiles + blockIdx.x];
        for (int ww = 0; ww < RS_THREADS / 32; ww++) {
            uint32_t c = cnt[ww][d];
            cnt[ww][d] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int r = 0; r < RS_ROUNDS; r++) {
        const size_t i = base + (size_t)r * 32 + lane;
        if (i < m) {
            const uint32_t pos = cnt[w][(k[r] >> shift) & 255u] + rank[r];
            keys_out[pos] = k[r];
            vals_out[pos] = v[r];
        }
    }
}

void radix_sort_stable_u32(WiringScratch& S, uint32_t* keys, uint32_t* vals, size_t m, int key_bits, uint32_t** keys_sorted,
                           uint32_t** vals_sorted, cudaStream_t st) {
    const size_t ntiles = (m + RS_TILE - 1) / RS_TILE;
    if (S.keys2.n < m) S.keys2.alloc(m);
    if (S.vals2.n < m) S.vals2.alloc(m);
    if (S.hist.n < 256 * ntiles) S.hist.alloc(256 * ntiles);
    if (S.offs.n < 256 * ntiles + 1) S.offs.alloc(256 * ntiles + 1);
    if (S.tile_sum.n < 256 * ntiles / 2048 + 2) S.tile_sum.alloc(256 * ntiles / 2048 + 2);
    uint32_t *ka = keys, *va = vals, *kb = S.keys2.p, *vb = S.vals2.p;
    for (int shift = 0; shift < key_bits; shift += 8) {
        ZP_LAUNCH(rs_hist_kernel, dim3((unsigned)ntiles), dim3(RS_THREADS), 0, st, ka, m, shift, ntiles, S.hist.p);
        u32_exclusive_scan(S.hist.p, S.offs.p, 256 * ntiles, S.tile_sum.p, st);
        ZP_LAUNCH(rs_scatter_kernel, dim3((unsigned)ntiles), dim3(RS_THREADS), 0, st, ka, va, m, shift, ntiles, S.offs.p, kb, vb);
        uint32_t* t = ka; ka = kb; kb = t;
        t = va; va = vb; vb = t;
    }
    *keys_sorted = ka;
    *vals_sorted = va;
}

// K_wire * omega_N^gate, K = 1, 7, 13, 17 by additions
ZP_D fr_t encode_cell(uint32_t cell, int logn, const fr_t* w_lo, const fr_t* w_hi) {
    const uint32_t gate = cell >> 2, wire = cell & 3u;
    uint32_t ex = gate << (NTT_LMAX - logn);
    uint32_t lo = ex & ((1u << NTT_LO_BITS) - 1), hi = ex >> NTT_LO_BITS;
    fr_t r = load_fr(&w_hi[hi]);
    if (lo) r = r * load_fr(&w_lo[lo]);
    if (wire == 0) return r;
    fr_t r2 = r.dbl(), r4 = r2.dbl(), r8 = r4.dbl();
    if (wire == 1) return r8 - r;
    if (wire == 2) return r8 + r4 + r;
    return r8.dbl() + r;
}

// identity permutation: sigma_k(omega^i) = K_k omega^i
__global__ void sigma_identity_kernel(fr_t* s0, fr_t* s1, fr_t* s2, fr_t* s3, int logn, const fr_t* w_lo, const fr_t* w_hi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> logn) return;
    fr_t* s[4] = {s0, s1, s2, s3};
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) store_fr(&s[k][i], encode_cell(((uint32_t)i << 2) | k, logn, w_lo, w_hi));
}
// first[var] = position of the variable's first entry in the grouped list
__global__ void group_first_kernel(const uint32_t* __restrict__ keys, size_t m, uint32_t* __restrict__ first) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    if (j == 0 || keys[j - 1] != keys[j]) first[keys[j]] = (uint32_t)j;
}
// sigma[cell j] = encode(cell of the next entry of the same variable, wrapping to the first)
__global__ void sigma_link_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ cells, size_t m,
                                  const uint32_t* __restrict__ first, fr_t* s0, fr_t* s1, fr_t* s2, fr_t* s3, int logn,
                                  const fr_t* w_lo, const fr_t* w_hi) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t key = keys[j], cur = cells[j];
    const uint32_t nxt = (j + 1 < m && keys[j + 1] == key) ? cells[j + 1] : cells[first[key]];
    fr_t* s[4] = {s0, s1, s2, s3};
    store_fr(&s[cur & 3u][cur >> 2], encode_cell(nxt, logn, w_lo, w_hi));
}

void sigma_from_wiring(WiringScratch& S, uint32_t* vars, uint32_t* cells, size_t m, uint32_t n_vars, int logn, const NttTables& T,
                       fr_t* const sigma[4], cudaStream_t st) {
    const size_t n = (size_t)1 << logn;
    ZP_LAUNCH(sigma_identity_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, sigma[0], sigma[1], sigma[2], sigma[3], logn,
              T.w_lo.p, T.w_hi.p);
    if (!m) return;
    int key_bits = 0;
    while (key_bits < 32 && ((uint64_t)1 << key_bits) < (uint64_t)n_vars) key_bits++;
    if (key_bits == 0) key_bits = 1;
    uint32_t *ks, *vs;
    radix_sort_stable_u32(S, vars, cells, m, key_bits, &ks, &vs, st);
    if (S.first.n < n_vars) S.first.alloc(n_vars);
    ZP_LAUNCH(group_first_kernel, dim3((unsigned)((m + 255) / 256)), dim3(256), 0, st, ks, m, S.first.p);
    ZP_LAUNCH(sigma_link_kernel, dim3((unsigned)((m + 255) / 256)), dim3(256), 0, st, ks, vs, m, S.first.p, sigma[0], sigma[1],
              sigma[2], sigma[3], logn, T.w_lo.p, T.w_hi.p);
}

}  // namespace zp
