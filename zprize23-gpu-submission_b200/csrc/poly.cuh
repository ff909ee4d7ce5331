// Element-wise, scan and evaluation kernels of the prover (host-side interface of poly.cu).
#pragma once
#include "common.cuh"
#include "ntt.cuh"
#include "host_math.hpp"
#include "msm.cuh"

namespace zp {

struct PolyScratch {
    DevBuf<fr_t> lv[8];   // partials of the chunked scans / Horner sweeps
    DevBuf<fr_t> tmp;
    DevBuf<uint32_t> flag;
    fr_t* host_pinned = nullptr;  // 64 KiB pinned staging for small D2H reads
    void init();
    ~PolyScratch();
};

// t = a + zeta b + zeta^2 c + zeta^3 d   (MultiSet::compress, lookup/multiset.rs:207-213)
void compress4(fr_t* out, const fr_t* a, const fr_t* b, const fr_t* c, const fr_t* d, const fr_t& zeta, size_t n, cudaStream_t st);
// query vector f (prover.rs:260-286): wires where q_lookup != 0, else (t[0],0,0,0), compressed with zeta
void query_f(fr_t* out, const fr_t* w0, const fr_t* w1, const fr_t* w2, const fr_t* w3, const fr_t* q_lookup, size_t n_real,
             const fr_t* t_ev, const fr_t& zeta, size_t n, cudaStream_t st);
// permutation ratio (permutation/mod.rs:629-752): num/den per gate, sigma given as evaluations on H
void perm_num_den(fr_t* num, fr_t* den, const fr_t* const w[4], const fr_t* const sigma[4], const fr_t& beta, const fr_t& gamma,
                  int logn, const NttTables& T, cudaStream_t st);
// lookup ratio (permutation/mod.rs:803-838)
void lookup_num_den(fr_t* num, fr_t* den, const fr_t* f, const fr_t* t, const fr_t* h1, const fr_t* h2, const fr_t& delta,
                    const fr_t& epsilon, size_t n, cudaStream_t st);
// den[i] <- num[i] / den[i]   (batched inversion, Montgomery trick per thread)
void ratio_inplace(const fr_t* num, fr_t* den, fr_t* scratch /* n elements */, size_t n, cudaStream_t st);
// out[0] = 1, out[i] = prod_{j<i} r[j]   (the grand product z(X) / z2(X); out may alias r)
void exclusive_prefix_product(PolyScratch& S, const fr_t* r, fr_t* out, size_t n, cudaStream_t st);
// p(z) for `count` (polynomial, point) pairs of n coefficients each; results land in host memory
void evaluate_many(PolyScratch& S, const fr_t* const* polys, const fr_t* points, int count, size_t n, fr_t* results_host,
                   cudaStream_t st);
// q = floor(p / (X - z)): q[i-1] = s_i, s_i = p_i + z s_{i+1}; q has n-1 coefficients, q[n-1] is set to 0
void divide_by_linear(PolyScratch& S, const fr_t* p, size_t n, const fr_t& z, fr_t* q, cudaStream_t st);
// out[i] = sum_k s[k] * polys[k][i]
void lincomb(fr_t* out, const fr_t* const* polys, const fr_t* scalars, int count, size_t n, cudaStream_t st);
// flag helpers (device reductions, result read by the host)
bool all_zero(PolyScratch& S, const fr_t* a, size_t n, cudaStream_t st);
bool all_equal_to_first(PolyScratch& S, const fr_t* a, size_t n, cudaStream_t st);
void fill(fr_t* out, const fr_t& v, size_t n, cudaStream_t st);

// h1, h2 = MultiSet::combine_split(t, f) on the device (lookup/multiset.rs:131-176): the multiset union of t and f
// ordered by first occurrence in t, split into its even- and odd-indexed halves.  Returns false when an element of f
// is not in t (Error::ElementNotIndexed).  The reference's native code skips this step (gen_proof.cuh:107-115).
struct CombineSplitScratch {
    DevBuf<uint32_t> table;   // open-addressing hash table: (first index, count) pairs
    DevBuf<uint32_t> cnt, offs, tile_sum, flag;
};
bool combine_split(CombineSplitScratch& S, const fr_t* t, const fr_t* f, size_t n, fr_t* h1, fr_t* h2, cudaStream_t st);
bool combine_split(CombineSplitScratch& S, const fr_t* t, size_t nt, const fr_t* f, size_t nf, fr_t* h1, fr_t* h2, cudaStream_t st);

struct QuotientArgs {
    int logn;                 // log2 N (the 8N coset has 2^(logn+3) points)
    const fr_t* w[4];         // wire coset evaluations (8N)
    const fr_t* z;            // z coset evaluations
    const fr_t* z2;           // nullptr when the lookup argument is trivial
    const fr_t *f, *table, *h1, *h2;
    // public input: PI(X) = pi_val * L_pos(X) and L_pos(x_i) = L_1(x_{i - 8 pos}) on the coset, so the PI stream is a
    // rotation of the resident l1 array (no NTT, no extra array); pi_count = 0 when the value is zero
    int pi_count;
    fr_t pi_val;
    uint32_t pi_shift;        // 8 * pos
    const fr_t* l1;           // L_1 on the coset (resident, depends on N only)
    const fr_t* sel[15];      // selector coset evaluations in ProverKeyC order; nullptr = identically zero
    const fr_t* sigma[4];
    fr_t alpha_sq;            // alpha^2 (hoisted out of the kernel)
    fr_t alpha, beta, gamma, delta, epsilon, zeta, range_sep, logic_sep, fixed_sep, var_sep, lookup_sep;
    fr_t vh_inv[8];           // 1 / (g^N w8^k - 1)
    fr_t coeff_d;             // JubJub d
    const fr_t *w_lo, *w_hi;  // omega tables (for x = g * omega_8N^i)
    fr_t beta_g;              // beta * 7 (coset generator folded into the permutation challenge)
    fr_t* out;
    size_t i_begin, i_count;  // index range of the coset handled by this launch (whole coset: 0, 8N)
    // coset_j >= 0: 2^coset_lc consecutive size-N cosets of the extended domain, coset_j + c (points g w_8N^(8 t + coset_j + c),
    // t < N; i_begin = 0, i_count = N << coset_lc).  w / z / z2 / f / table / h1 / h2 and out are COMPACT arrays indexed by
    // c * N + t ("next" = t + 1 inside the coset); the prover-key streams and l1 stay in natural 8N order and are read at
    // 8 t + coset_j + c, so neighbouring threads (c fastest) read neighbouring elements.  -1: the whole 8N domain, natural order.
    int coset_j, coset_lc;
    // 1: sel / sigma / l1 are this coset's COMPACT copies indexed by t (one coset per launch, coset_lc = 0); the public-input
    // rotation of l1 stays inside the coset (L_pos(x_{8 t + j}) = L_1(x_{8 (t - pos) + j}))
    int compact_key = 0;
};
void quotient_evals(const QuotientArgs& a, cudaStream_t st);

}  // namespace zp
