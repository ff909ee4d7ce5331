// Fr NTT / iNTT / coset-NTT / coset-iNTT, natural order in and out.
// Host-side interface of ntt.cu.  Replaces the reference's `Ntt/Intt/Ntt_coset/Intt_coset::forward`
// ("Prize 1B/plonk-core/lib/PLONK/utils/function.cu":249-273 -> "…/utils/zkp/cuda/zksnark_ntt.cu":16-92
// -> sppark `_CT_NTT`, `bit_rev_permutation`, `LDE_distribute_powers`); semantics of ark-poly 0.3
// `GeneralEvaluationDomain::{fft, ifft, coset_fft, coset_ifft}`.
#pragma once
#include "common.cuh"
#include <map>

namespace zp {

enum NttKind { NTT_FWD = 0, NTT_INV = 1, NTT_COSET_FWD = 2, NTT_COSET_INV = 3 };

static const int NTT_LMAX = 26;      // largest supported log2 domain (2^25 = 8N at HEIGHT=15)
static const int NTT_LO_BITS = 13;   // two-level power tables: x^e = hi[e >> 13] * lo[e & 8191]

struct NttTables {
    // all tables resident for the life of the context (built once, not per transform object as the
    // reference does 9 times per proof — SURVEY §2.3 `generate_*_twiddles`)
    DevBuf<fr_t> w_lo, w_hi;        // omega_{2^LMAX}^e
    DevBuf<fr_t> g_lo, g_hi;        // 7^e
    DevBuf<fr_t> gi_lo, gi_hi;      // 7^-e
    fr_t ninv[NTT_LMAX + 1];        // 2^-k (host copies, passed by value to kernels)
    fr_t omega[NTT_LMAX + 1];       // primitive 2^k-th roots (host)
    fr_t omega_inv[NTT_LMAX + 1];
    bool ready = false;
    void init(cudaStream_t st);
    // Direct per-pass tables (built on first use, resident afterwards): the inter-pass twiddle omega_N^{n * ks} of pass p as
    // tw[(n << lk) | ks], and for the coset iNTT the output factor 2^-logn * 7^-pos as out[pos].  One load + ONE product per
    // element instead of the two-level lookup's two (three for the coset output); costs 32 B of extra HBM read per element on
    // a pass that is multiplier-bound.  Tables above 2^tw_max_log entries are not built (two-level lookup is used instead).
    mutable std::map<uint32_t, DevBuf<fr_t>> direct;
    int tw_max_log = 25;  // ZP_NTT_TW_MAX_LOG; 0 disables
    int tw_min_log = 16;  // transforms below 2^tw_min_log keep the two-level lookup (ZP_NTT_TW_MIN_LOG)
    const fr_t* pass_table(int logn, int inverse, int lr, int lk, cudaStream_t st) const;
    const fr_t* coset_out_table(int logn, cudaStream_t st) const;
};

struct NttScratch {
    DevBuf<fr_t> s1, s2;
    void reserve(size_t n) {
        if (s1.n < n) s1.alloc(n);
        if (s2.n < n) s2.alloc(n);
    }
};

// out[0..2^logn) = transform(in[0..n_in) zero-padded to 2^logn).  in == out is allowed.
void ntt_run(const NttTables& T, NttScratch& S, NttKind kind, int logn, const fr_t* in, size_t n_in, fr_t* out,
             cudaStream_t st);

// Coset-by-coset view of the size-8N extended domain (used when the quotient round is sharded over GPUs): the points
// g w_8N^(8 i + j), i < N, form the coset (g w_8N^j) H_N, so the evaluations of a degree < N polynomial on it are ONE size-N
// coset NTT of the coefficients pre-multiplied by w_8N^(j m).
//   ntt_coset_shift : out[m] = in[m] * w_{2^logn_big}^(+- j m), m < n   (in == out allowed)
//   ntt_combine8    : given P_j = size-N coset-iNTT over coset j of the quotient values (j < 8, PJ[j * N + m]), writes the
//                     8N coefficients t[m' * N + m] = g^(-N m') / 8 * sum_j w_8^(-j m') P_j[m]  (the split t_1 .. t_8)
void ntt_coset_shift(const NttTables& T, const fr_t* in, fr_t* out, size_t n, int logn_big, int j, bool inverse, cudaStream_t st);
void ntt_combine8(const NttTables& T, const fr_t* PJ, fr_t* t_out, int logn, cudaStream_t st);

// host-side field helpers shared by the driver
fr_t fr_two_adic_root_host();
fr_t fr_generator_host();

}  // namespace zp

// ---- four-step NTT sharded over G ranks (one process per GPU) -----------------------------------
// Rank r holds the contiguous block x[r*M, (r+1)*M), M = N/G, and ends with the contiguous block X[r*M, (r+1)*M) of
// the transform (natural order, same semantics as ntt_run).  `alltoall(user, send, recv, bytes_per_peer)` exchanges
// device memory: chunk p of `send` goes to rank p, chunk q of `recv` comes from rank q (NCCL all-to-all over NVLink).
namespace zp {
typedef int (*ntt_alltoall_fn)(void* user, const void* send_dev, void* recv_dev, size_t bytes_per_peer);
void ntt_sharded_run(const NttTables& T, NttScratch& S, NttKind kind, int logn_total, int rank, int world, const fr_t* in_local,
                     fr_t* out_local, fr_t* tmp_a, fr_t* tmp_b, ntt_alltoall_fn a2a, void* user, cudaStream_t st);
}  // namespace zp
