// Fr NTT family for sm_100a — see ntt.cuh for the contract and DESIGN.md §NTT for the derivation.
//
// Decomposition: N = R_0 R_1 … R_{P-1} (R_p = 2^{lr_p}, lr_p <= 9).  Before pass p the array is the
// tensor [n_p][n_{p+1}]…[n_{P-1}][k_{p-1}]…[k_0] (row-major, leftmost slowest); pass p transforms the
// leading digit n_p -> k_p and writes [n_{p+1}]…[n_{P-1}][k_p][k_{p-1}]…[k_0].  p = 0 is the natural
// input order, after the last pass the layout is [k_{P-1}]…[k_0] = natural output order: no separate
// bit-reversal pass, no separate coset-power pass, no separate 1/N pass, zero padding is implicit.
// Each CTA owns a tile of R_p x C elements (C consecutive values of the flattened trailing index q), so
// every global read is a C*32-byte contiguous segment and every write a >= C*32-byte one.
// Inside the tile the size-R_p transforms run as radix-2 DIF stages in shared memory (output index read
// back bit-reversed).  The inter-pass twiddle  omega_N^{n_p * S_p * K_{p-1}}  is applied on load.
#include "ntt.cuh"

namespace zp {

// TWO_ADIC_ROOT_OF_UNITY and multiplicative generator 7 of BLS12-381 Fr, Montgomery form
// (same values as "Prize 1B/plonk-core/lib/PLONK/src/bls12_381/fr.cuh":39-53).
fr_t fr_two_adic_root_host() {
    fr_t r;
    const uint64_t v[4] = {13381757501831005802ULL, 6564924994866501612ULL, 789602057691799140ULL, 6625830629041353339ULL};
    memcpy(r.l, v, 32);
    return r;
}
fr_t fr_generator_host() {
    fr_t r;
    const uint64_t v[4] = {64424509425ULL, 1721329240476523535ULL, 18418692815241631664ULL, 3824455624000121028ULL};
    memcpy(r.l, v, 32);
    return r;
}

// table[i] = base^(i << shift)
__global__ void power_table_kernel(fr_t* table, fr_t base, int shift, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t e = (uint64_t)i << shift;
    store_fr(&table[i], base.pow_u64(e));
}

void NttTables::init(cudaStream_t st) {
    if (ready) return;
    fr_t w = fr_two_adic_root_host();  // order 2^32
    for (int i = NTT_LMAX; i < 32; i++) w = w.sqr();
    omega[NTT_LMAX] = w;
    for (int k = NTT_LMAX - 1; k >= 0; k--) omega[k] = omega[k + 1].sqr();
    fr_t two_inv = fr_t::from_u32(2).inverse();
    ninv[0] = fr_t::one();
    for (int k = 1; k <= NTT_LMAX; k++) ninv[k] = ninv[k - 1] * two_inv;
    for (int k = 0; k <= NTT_LMAX; k++) omega_inv[k] = omega[k].inverse();
    const int LO = 1 << NTT_LO_BITS, HI = 1 << (NTT_LMAX - NTT_LO_BITS);
    w_lo.alloc(LO); w_hi.alloc(HI);
    g_lo.alloc(LO); g_hi.alloc(HI);
    gi_lo.alloc(LO); gi_hi.alloc(HI);
    fr_t g = fr_generator_host(), gi = g.inverse();
    ZP_LAUNCH(power_table_kernel, dim3((LO + 255) / 256), dim3(256), 0, st, w_lo.p, omega[NTT_LMAX], 0, LO);
    ZP_LAUNCH(power_table_kernel, dim3((HI + 255) / 256), dim3(256), 0, st, w_hi.p, omega[NTT_LMAX], NTT_LO_BITS, HI);
    ZP_LAUNCH(power_table_kernel, dim3((LO + 255) / 256), dim3(256), 0, st, g_lo.p, g, 0, LO);
    ZP_LAUNCH(power_table_kernel, dim3((HI + 255) / 256), dim3(256), 0, st, g_hi.p, g, NTT_LO_BITS, HI);
    ZP_LAUNCH(power_table_kernel, dim3((LO + 255) / 256), dim3(256), 0, st, gi_lo.p, gi, 0, LO);
    ZP_LAUNCH(power_table_kernel, dim3((HI + 255) / 256), dim3(256), 0, st, gi_hi.p, gi, NTT_LO_BITS, HI);
    if (const char* e = getenv("ZP_NTT_TW_MAX_LOG")) tw_max_log = atoi(e);
    if (const char* e = getenv("ZP_NTT_TW_MIN_LOG")) tw_min_log = atoi(e);
    ready = true;
}

struct NttPassParams {
    int logn;       // log2 N
    int lr;         // log2 R_p
    int lc;         // log2 C (tile width)
    int lk;         // log2 K_done = lr_0 + … + lr_{p-1}
    int first, last;
    int inverse;    // use omega^-1
    int coset;      // 0 none, 1 multiply input by g^n (first pass), 2 multiply output by g^-k (last pass)
    size_t n_in;    // elements >= n_in of the input are implicit zeros (first pass only)
    const fr_t *w_lo, *w_hi, *c_lo, *c_hi;
    const fr_t* tw;      // direct inter-pass twiddle table [(n << lk) | ks] of this pass, or null (two-level lookup)
    const fr_t* out_tw;  // direct output factors 2^-logn * 7^-pos of the coset iNTT, or null
    fr_t ninv;      // 2^-logn (used when inverse && last)
};

// omega_{2^LMAX}^e, e in [0, 2^LMAX)
ZP_D fr_t tw_lookup(const fr_t* lo_t, const fr_t* hi_t, uint32_t e) {
    uint32_t lo = e & ((1u << NTT_LO_BITS) - 1), hi = e >> NTT_LO_BITS;
    if (lo == 0) return load_fr(&hi_t[hi]);
    fr_t a = load_fr(&lo_t[lo]);
    if (hi == 0) return a;
    return a * load_fr(&hi_t[hi]);
}

ZP_D void sm_store(uint4* sm, int half_stride, int idx, const fr_t& v) {
    sm[idx] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    sm[half_stride + idx] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
ZP_D fr_t sm_load(const uint4* sm, int half_stride, int idx) {
    uint4 a = sm[idx], b = sm[half_stride + idx];
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

__global__ void __launch_bounds__(1024) ntt_pass_kernel(const fr_t* __restrict__ in, fr_t* __restrict__ out, NttPassParams p) {
    ZP_DYN_SMEM(uint4, sm);
    const int R = 1 << p.lr, C = 1 << p.lc, RC = R * C;
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t q0 = (size_t)blockIdx.x << p.lc;
    const uint32_t emask = (1u << NTT_LMAX) - 1;

    // ---- load tile [n][c], applying coset powers / inter-pass twiddles
    for (int e = tid; e < RC; e += nt) {
        int n = e >> p.lc, c = e & (C - 1);
        size_t q = q0 + c;
        size_t pos = ((size_t)n << (p.logn - p.lr)) + q;
        fr_t v;
        bool nz = true;
        if (p.first) {
            nz = pos < p.n_in;
            v = nz ? load_fr(&in[pos]) : fr_t::zero();
            if (nz && p.coset == 1 && pos != 0) {
                uint32_t lo = (uint32_t)pos & ((1u << NTT_LO_BITS) - 1), hi = (uint32_t)(pos >> NTT_LO_BITS);
                fr_t f = load_fr(&p.c_lo[lo]);
                if (hi) f = f * load_fr(&p.c_hi[hi]);
                v = v * f;
            }
        } else {
            v = load_fr(&in[pos]);
            uint32_t ks = (uint32_t)q & ((1u << p.lk) - 1);
            uint32_t ex = ((uint32_t)n * ks) << (NTT_LMAX - (p.lk + p.lr));  // theta = omega_{T_{p+1}}
            if (ex) {
                if (p.tw) {
                    v = v * load_fr(&p.tw[((size_t)n << p.lk) | ks]);
                } else {
                    if (p.inverse) ex = ((1u << NTT_LMAX) - ex) & emask;
                    v = v * tw_lookup(p.w_lo, p.w_hi, ex);
                }
            }
        }
        sm_store(sm, RC, e, v);
    }
    __syncthreads();

    // ---- radix-2 DIF stages over n (in place, result index bit-reversed)
    const int nb = RC >> 1;
    for (int s = 0; s < p.lr; s++) {
        const int lhalf = p.lr - 1 - s;
        const int half = 1 << lhalf;
        for (int b = tid; b < nb; b += nt) {
            int c = b & (C - 1), t = b >> p.lc;
            int j = t & (half - 1), blk = t >> lhalf;
            int i0 = ((blk << (lhalf + 1)) + j) * C + c, i1 = i0 + half * C;
            fr_t u = sm_load(sm, RC, i0), v = sm_load(sm, RC, i1);
            sm_store(sm, RC, i0, u + v);
            fr_t d = u - v;
            if (j) {
                uint32_t ex = ((uint32_t)j << s) << (NTT_LMAX - p.lr);
                if (p.inverse) ex = ((1u << NTT_LMAX) - ex) & emask;
                d = d * tw_lookup(p.w_lo, p.w_hi, ex);
            }
            sm_store(sm, RC, i1, d);
        }
        __syncthreads();
    }

    // ---- store: X[k] sits at bit-reversed row
    const int sh = 32 - p.lr;
    for (int e = tid; e < RC; e += nt) {
        int k, c;
        size_t pos;
        if (p.first) {  // K_done = 1: whole tile is one contiguous block [c][k]
            k = e & (R - 1);
            c = e >> p.lr;
            pos = (q0 << p.lr) + e;
        } else {
            k = e >> p.lc;
            c = e & (C - 1);
            size_t q = q0 + c;
            size_t rest = q >> p.lk, ks = q & (((size_t)1 << p.lk) - 1);
            pos = (rest << (p.lk + p.lr)) + ((size_t)k << p.lk) + ks;
        }
        int row = p.lr ? (int)(__brev((uint32_t)k) >> sh) : 0;
        fr_t v = sm_load(sm, RC, row * C + c);
        if (p.last) {
            if (p.coset == 2 && p.out_tw) {
                v = v * load_fr(&p.out_tw[pos]);
            } else if (p.coset == 2) {
                uint32_t lo = (uint32_t)pos & ((1u << NTT_LO_BITS) - 1), hi = (uint32_t)(pos >> NTT_LO_BITS);
                fr_t f = p.ninv;
                if (lo) f = f * load_fr(&p.c_lo[lo]);
                if (hi) f = f * load_fr(&p.c_hi[hi]);
                v = v * f;
            } else if (p.inverse) {
                v = v * p.ninv;
            }
        }
        store_fr(&out[pos], v);
    }
}

// ---- register-resident variant of the pass kernel (lr >= 6, 1024-element tile, 128 threads x 8 elements) -----------
// The lr radix-2 DIF stages of a tile are grouped 3 + 3 + (lr - 6): in every group a thread holds the 8 elements that
// differ in the group's three row bits and runs the three stages in REGISTERS (radix-8 butterfly network, 7 twiddle
// loads), so a tile needs 3 shared-memory exchanges and 3 barriers instead of lr (<= 9) of each; the first group is fused
// with the global load (coset power / zero padding / inter-pass twiddle applied in flight), the last group's twiddles are
// the constants w_8^j.  Rows are padded by one row per eight ((row + row / 8) * C + c) so that all three access patterns
// hit distinct banks.  Same arithmetic, same operand order as ntt_pass_kernel: results are bit-identical.
ZP_D int pad_idx(int row, int c, int lc) { return ((row + (row >> 3)) << lc) + c; }

// x[i1] = (x[i0] - x[i1]) * w^ex, x[i0] += x[i1]  with the twiddle omega_{2^LMAX}^(+-ex) looked up unless ex == 0
ZP_D void bfly(fr_t& a, fr_t& b, uint32_t ex, const NttPassParams& p) {
    fr_t u = a, v = b;
    a = u + v;
    fr_t d = u - v;
    if (ex) {
        if (p.inverse) ex = ((1u << NTT_LMAX) - ex) & ((1u << NTT_LMAX) - 1);
        d = d * tw_lookup(p.w_lo, p.w_hi, ex);
    }
    b = d;
}
// three DIF stages on x[0..8) for the row bits (hi, mid, lo) = register index bits (2, 1, 0); j0 = the part of the in-tile
// index below the group's bits, unit = weight of the group's lowest bit, s0 = index of the group's first stage
ZP_D void dif3(fr_t x[8], uint32_t j0, uint32_t unit, int s0, int sh, const NttPassParams& p) {
#pragma unroll
    for (int a = 0; a < 4; a++) bfly(x[a], x[a + 4], ((a * unit + j0) << s0) << sh, p);
#pragma unroll
    for (int a = 0; a < 2; a++) {
        const uint32_t ex = ((a * unit + j0) << (s0 + 1)) << sh;
        bfly(x[a], x[a + 2], ex, p);
        bfly(x[a + 4], x[a + 6], ex, p);
    }
    {
        const uint32_t ex = (j0 << (s0 + 2)) << sh;
#pragma unroll
        for (int a = 0; a < 8; a += 2) bfly(x[a], x[a + 1], ex, p);
    }
}

// one DIF stage over register pairs (e, e + HALF); twiddle exponent ((e mod HALF) << s) << sh (compile-time register indices)
template <int HALF>
ZP_D void dif_stage_regs(fr_t x[8], int s, int sh, const NttPassParams& p) {
#pragma unroll
    for (int e = 0; e < 8; e++)
        if (!(e & HALF)) bfly(x[e], x[e + HALF], (((uint32_t)e & (HALF - 1)) << s) << sh, p);
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) ntt_pass8_kernel(const fr_t* __restrict__ in, fr_t* __restrict__ out, NttPassParams p) {
    ZP_DYN_SMEM(uint4, sm);
    const int R = 1 << p.lr, C = 1 << p.lc, RC = R * C;  // RC == 1024
    const int RCp = (R + (R >> 3)) * C;                   // padded element count
    const int tid = threadIdx.x;
    const size_t q0 = (size_t)blockIdx.x << p.lc;
    const int sh = NTT_LMAX - p.lr;
    const int c = tid & (C - 1);
    fr_t x[8];

    // ---- group 1 (row bits lr-1 .. lr-3) fused with the load: thread (rho, c) holds rows a * R/8 + rho
    {
        const int rho = tid >> p.lc;
        const size_t q = q0 + c;
#pragma unroll
        for (int a = 0; a < 8; a++) {
            const int n = a * (R >> 3) + rho;
            const size_t pos = ((size_t)n << (p.logn - p.lr)) + q;
            fr_t v;
            if (p.first) {
                const bool nz = pos < p.n_in;
                v = nz ? load_fr(&in[pos]) : fr_t::zero();
                if (nz && p.coset == 1 && pos != 0) {
                    uint32_t lo = (uint32_t)pos & ((1u << NTT_LO_BITS) - 1), hi = (uint32_t)(pos >> NTT_LO_BITS);
                    fr_t f = load_fr(&p.c_lo[lo]);
                    if (hi) f = f * load_fr(&p.c_hi[hi]);
                    v = v * f;
                }
            } else {
                v = load_fr(&in[pos]);
                uint32_t ks = (uint32_t)q & ((1u << p.lk) - 1);
                uint32_t ex = ((uint32_t)n * ks) << (NTT_LMAX - (p.lk + p.lr));
                if (ex) {
                    if (p.tw) {
                        v = v * load_fr(&p.tw[((size_t)n << p.lk) | ks]);
                    } else {
                        if (p.inverse) ex = ((1u << NTT_LMAX) - ex) & ((1u << NTT_LMAX) - 1);
                        v = v * tw_lookup(p.w_lo, p.w_hi, ex);
                    }
                }
            }
            x[a] = v;
        }
        dif3(x, (uint32_t)rho, (uint32_t)(R >> 3), 0, sh, p);
#pragma unroll
        for (int a = 0; a < 8; a++) sm_store(sm, RCp, pad_idx(a * (R >> 3) + rho, c, p.lc), x[a]);
    }
    __syncthreads();

    // ---- group 2 (row bits lr-4 .. lr-6): thread (a, r, c) holds rows a * R/8 + b * R/64 + r
    {
        const int u = tid >> p.lc;  // (a, r)
        const int r = u & ((R >> 6) - 1), a = u >> (p.lr - 6);
        const int row0 = a * (R >> 3) + r;
#pragma unroll
        for (int b = 0; b < 8; b++) x[b] = sm_load(sm, RCp, pad_idx(row0 + b * (R >> 6), c, p.lc));
        dif3(x, (uint32_t)r, (uint32_t)(R >> 6), 3, sh, p);
#pragma unroll
        for (int b = 0; b < 8; b++) sm_store(sm, RCp, pad_idx(row0 + b * (R >> 6), c, p.lc), x[b]);
    }
    __syncthreads();

    // ---- group 3 (the lr - 6 lowest row bits): thread (u, c) holds rows 8 u .. 8 u + 7; twiddles are w_8^j / w_4^j
    const int m = p.lr - 6;
    if (m > 0) {
        const int u = tid >> p.lc;
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = sm_load(sm, RCp, pad_idx(8 * u + e, c, p.lc));
        if (m == 3) {
            dif_stage_regs<4>(x, 6, sh, p);
            dif_stage_regs<2>(x, 7, sh, p);
            dif_stage_regs<1>(x, 8, sh, p);
        } else if (m == 2) {
            dif_stage_regs<2>(x, 6, sh, p);
            dif_stage_regs<1>(x, 7, sh, p);
        } else {
            dif_stage_regs<1>(x, 6, sh, p);
        }
#pragma unroll
        for (int e = 0; e < 8; e++) sm_store(sm, RCp, pad_idx(8 * u + e, c, p.lc), x[e]);
        __syncthreads();
    }

    // ---- store: X[k] sits at the bit-reversed row (identical to ntt_pass_kernel)
    const int shb = 32 - p.lr;
    for (int e = tid; e < RC; e += 128) {
        int k, cc;
        size_t pos;
        if (p.first) {
            k = e & (R - 1);
            cc = e >> p.lr;
            pos = (q0 << p.lr) + e;
        } else {
            k = e >> p.lc;
            cc = e & (C - 1);
            size_t q = q0 + cc;
            size_t rest = q >> p.lk, ks = q & (((size_t)1 << p.lk) - 1);
            pos = (rest << (p.lk + p.lr)) + ((size_t)k << p.lk) + ks;
        }
        int row = (int)(__brev((uint32_t)k) >> shb);
        fr_t v = sm_load(sm, RCp, pad_idx(row, cc, p.lc));
        if (p.last) {
            if (p.coset == 2 && p.out_tw) {
                v = v * load_fr(&p.out_tw[pos]);
            } else if (p.coset == 2) {
                uint32_t lo = (uint32_t)pos & ((1u << NTT_LO_BITS) - 1), hi = (uint32_t)(pos >> NTT_LO_BITS);
                fr_t f = p.ninv;
                if (lo) f = f * load_fr(&p.c_lo[lo]);
                if (hi) f = f * load_fr(&p.c_hi[hi]);
                v = v * f;
            } else if (p.inverse) {
                v = v * p.ninv;
            }
        }
        store_fr(&out[pos], v);
    }
}

// tw[(n << lk) | ks] = omega_{2^(lk+lr)}^{+- n * ks}
__global__ void ntt_pass_table_kernel(fr_t* __restrict__ tw, int lr, int lk, int inverse, const fr_t* __restrict__ w_lo,
                                      const fr_t* __restrict__ w_hi) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >> (lr + lk)) return;
    uint32_t n = (uint32_t)(idx >> lk), ks = (uint32_t)idx & ((1u << lk) - 1);
    uint32_t ex = (n * ks) << (NTT_LMAX - (lk + lr));
    if (inverse) ex = ((1u << NTT_LMAX) - ex) & ((1u << NTT_LMAX) - 1);
    store_fr(&tw[idx], ex ? tw_lookup(w_lo, w_hi, ex) : fr_t::one());
}
// out[pos] = ninv * 7^-pos
__global__ void ntt_coset_out_table_kernel(fr_t* __restrict__ out, size_t n, fr_t ninv, const fr_t* __restrict__ c_lo,
                                           const fr_t* __restrict__ c_hi) {
    size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    uint32_t lo = (uint32_t)pos & ((1u << NTT_LO_BITS) - 1), hi = (uint32_t)(pos >> NTT_LO_BITS);
    fr_t f = ninv;
    if (lo) f = f * load_fr(&c_lo[lo]);
    if (hi) f = f * load_fr(&c_hi[hi]);
    store_fr(&out[pos], f);
}
const fr_t* NttTables::pass_table(int logn, int inverse, int lr, int lk, cudaStream_t st) const {
    if (lr + lk > tw_max_log) return nullptr;
    // the table depends on (lr, lk, direction) only: omega_{2^(lk+lr)}
    uint32_t key = (uint32_t)lr | ((uint32_t)lk << 8) | ((uint32_t)(inverse ? 1 : 0) << 16);
    (void)logn;
    auto it = direct.find(key);
    if (it == direct.end()) {
        DevBuf<fr_t> b((size_t)1 << (lr + lk));
        size_t cnt = (size_t)1 << (lr + lk);
        ZP_LAUNCH(ntt_pass_table_kernel, dim3((unsigned)((cnt + 255) / 256)), dim3(256), 0, st, b.p, lr, lk, inverse, w_lo.p, w_hi.p);
        // built once per context; transforms run on two streams (Prover::st / st2), so the table is complete before ANY
        // stream can see its pointer
        ZP_CUDA(cudaStreamSynchronize(st));
        it = direct.emplace(key, std::move(b)).first;
    }
    return it->second.p;
}
const fr_t* NttTables::coset_out_table(int logn, cudaStream_t st) const {
    if (logn > tw_max_log) return nullptr;
    uint32_t key = 0x80000000u | (uint32_t)logn;
    auto it = direct.find(key);
    if (it == direct.end()) {
        size_t cnt = (size_t)1 << logn;
        DevBuf<fr_t> b(cnt);
        ZP_LAUNCH(ntt_coset_out_table_kernel, dim3((unsigned)((cnt + 255) / 256)), dim3(256), 0, st, b.p, cnt, ninv[logn], gi_lo.p,
                  gi_hi.p);
        ZP_CUDA(cudaStreamSynchronize(st));
        it = direct.emplace(key, std::move(b)).first;
    }
    return it->second.p;
}

void ntt_run(const NttTables& T, NttScratch& S, NttKind kind, int logn, const fr_t* in, size_t n_in, fr_t* out,
             cudaStream_t st) {
    if (logn > NTT_LMAX) throw std::runtime_error("ntt_run: domain larger than 2^26");
    const size_t N = (size_t)1 << logn;
    if (n_in > N) n_in = N;
    if (logn == 0) {
        if (n_in) ZP_CUDA(cudaMemcpyAsync(out, in, sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
        else ZP_CUDA(cudaMemsetAsync(out, 0, sizeof(fr_t), st));
        return;
    }
    static const int KMAX = getenv("ZP_NTT_KMAX") ? atoi(getenv("ZP_NTT_KMAX")) : 9;
    int P = (logn + KMAX - 1) / KMAX;
    int bits[8];
    for (int p = 0; p < P; p++) bits[p] = logn / P + (p < logn % P ? 1 : 0);
    bool inverse = (kind == NTT_INV || kind == NTT_COSET_INV);
    if (P > 1) S.reserve(N);
    const fr_t* src = in;
    int lk = 0;
    for (int p = 0; p < P; p++) {
        bool last = (p == P - 1);
        fr_t* dst;
        if (last) {
            dst = out;
            if (dst == src) {  // single pass, in == out: go through scratch
                S.reserve(N);
                dst = S.s1.p;
            }
        } else {
            dst = (p & 1) ? S.s2.p : S.s1.p;
        }
        NttPassParams pp;
        pp.logn = logn;
        pp.lr = bits[p];
        int lq = logn - bits[p];  // log2 of number of q values
        static int tile_log = getenv("ZP_NTT_TILE_LOG") ? atoi(getenv("ZP_NTT_TILE_LOG")) : 10;
        static int max_threads = getenv("ZP_NTT_THREADS") ? atoi(getenv("ZP_NTT_THREADS")) : 128;
        // tile of 2^tile_log elements.  Measured on B200 (profiles/r01_ntt_tile_sweep.log): 1024-element (32 KiB) tiles
        // with 128 threads beat 4096-element tiles with 512 threads by 19 % (2^25 coset NTT 13.2 -> 10.7 ms): seven
        // resident CTAs per SM hide the barrier + multiplier latency better than one big CTA.
        int lc = tile_log - bits[p];
        if (lc < 0) lc = 0;
        if (lc > lq) lc = lq;
        if (p > 0 && lc > lk) lc = lk;  // C must divide K_done
        if (lc < 0) lc = 0;
        pp.lc = lc;
        pp.lk = lk;
        pp.first = (p == 0);
        pp.last = last;
        pp.inverse = inverse;
        pp.coset = (kind == NTT_COSET_FWD) ? 1 : (kind == NTT_COSET_INV ? 2 : 0);
        pp.n_in = n_in;
        pp.w_lo = T.w_lo.p;
        pp.w_hi = T.w_hi.p;
        pp.c_lo = (kind == NTT_COSET_FWD) ? T.g_lo.p : T.gi_lo.p;
        pp.c_hi = (kind == NTT_COSET_FWD) ? T.g_hi.p : T.gi_hi.p;
        pp.ninv = T.ninv[logn];
        pp.tw = (p > 0 && logn >= T.tw_min_log) ? T.pass_table(logn, inverse ? 1 : 0, pp.lr, lk, st) : nullptr;
        pp.out_tw = (last && kind == NTT_COSET_INV && logn >= T.tw_min_log) ? T.coset_out_table(logn, st) : nullptr;
        int RC = 1 << (pp.lr + pp.lc);
        size_t smem = (size_t)RC * 32;
        int threads = RC / 2 < max_threads ? (RC / 2 < 32 ? 32 : RC / 2) : max_threads;
        unsigned grid = (unsigned)(((size_t)1 << lq) >> lc);
#ifndef ZP_EMU
        static bool attr_set = false;
        if (!attr_set) {
            ZP_CUDA(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            // 64 KiB tiles: ask for the full shared-memory carve-out so that 3 CTAs are resident per SM
            ZP_CUDA(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            attr_set = true;
        }
#endif
        // ZP_NTT_REG = 4 | 5 | 6: the register-resident radix-8 kernel at that many resident CTAs per SM (128 / 102 / 85
        // registers).  Measured SLOWER than the shared-memory radix-2 kernel at every setting (profiles/r02k_ntt_sweep.log),
        // so it is off by default; kept for the record, parity-tested under ZP_NTT_REG=4.
        const char* reg_env = getenv("ZP_NTT_REG");  // read per call: the tests switch it on for single transforms
        const int use_reg = reg_env ? atoi(reg_env) : 0;
        if (use_reg && pp.lr >= 6 && RC == 1024 && threads == 128) {
            const size_t smem8 = (size_t)((1 << pp.lr) + (1 << (pp.lr - 3))) * (1 << pp.lc) * 32;
            auto launch8 = [&](auto kern) {
#ifndef ZP_EMU
                ZP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
                ZP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
#endif
                ZP_LAUNCH(kern, dim3(grid), dim3(128), smem8, st, src, dst, pp);
            };
            if (use_reg == 6) launch8(ntt_pass8_kernel<6>);
            else if (use_reg == 5) launch8(ntt_pass8_kernel<5>);
            else launch8(ntt_pass8_kernel<4>);
        } else {
            ZP_LAUNCH(ntt_pass_kernel, dim3(grid), dim3(threads), smem, st, src, dst, pp);
        }
        if (last && dst != out) ZP_CUDA(cudaMemcpyAsync(out, dst, N * sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
        src = dst;
        lk += bits[p];
    }
}

}  // namespace zp

namespace zp {

__global__ void __launch_bounds__(256) ntt_coset_shift_kernel(const fr_t* __restrict__ in, fr_t* __restrict__ out, size_t n,
                                                              int logn_big, uint32_t j, int inverse,
                                                              const fr_t* __restrict__ w_lo, const fr_t* __restrict__ w_hi) {
    size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    fr_t v = load_fr(&in[m]);
    uint32_t e = (uint32_t)(((uint64_t)j * m) & (((uint64_t)1 << logn_big) - 1));
    uint32_t ex = e << (NTT_LMAX - logn_big);
    if (ex) {
        if (inverse) ex = ((1u << NTT_LMAX) - ex) & ((1u << NTT_LMAX) - 1);
        v = v * tw_lookup(w_lo, w_hi, ex);
    }
    store_fr(&out[m], v);
}
void ntt_coset_shift(const NttTables& T, const fr_t* in, fr_t* out, size_t n, int logn_big, int j, bool inverse, cudaStream_t st) {
    if (!n) return;
    ZP_LAUNCH(ntt_coset_shift_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, in, out, n, logn_big, (uint32_t)j,
              inverse ? 1 : 0, T.w_lo.p, T.w_hi.p);
}

struct Combine8Consts {
    fr_t w[4];      // w_8^-e, e < 4
    fr_t scale[8];  // g^(-N m') / 8
};
// size-8 inverse DFT across the cosets (radix-2 DIF, result index bit-reversed), one thread per coefficient index m
__global__ void __launch_bounds__(128) ntt_combine8_kernel(const fr_t* __restrict__ PJ, fr_t* __restrict__ t_out, size_t n,
                                                           Combine8Consts c) {
    size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    fr_t x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = load_fr(&PJ[(size_t)j * n + m]);
#pragma unroll
    for (int s = 0; s < 3; s++) {
        const int half = 4 >> s;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int jj = b & (half - 1), blk = b / half;
            const int i0 = blk * 2 * half + jj, i1 = i0 + half;
            fr_t u = x[i0], v = x[i1];
            x[i0] = u + v;
            fr_t d = u - v;
            const int e = jj << s;
            x[i1] = e ? d * c.w[e] : d;
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int r = ((k & 1) << 2) | (k & 2) | ((k & 4) >> 2);  // X[k] sits at the bit-reversed position
        store_fr(&t_out[(size_t)k * n + m], x[r] * c.scale[k]);
    }
}
void ntt_combine8(const NttTables& T, const fr_t* PJ, fr_t* t_out, int logn, cudaStream_t st) {
    const size_t n = (size_t)1 << logn;
    Combine8Consts c;
    c.w[0] = fr_t::one();
    for (int e = 1; e < 4; e++) c.w[e] = c.w[e - 1] * T.omega_inv[3];
    // g^(-N): 7^-1 squared logn times
    fr_t gin = fr_generator_host().inverse();
    for (int k = 0; k < logn; k++) gin = gin.sqr();
    fr_t sc = T.ninv[3];
    for (int k = 0; k < 8; k++) {
        c.scale[k] = sc;
        sc = sc * gin;
    }
    ZP_LAUNCH(ntt_combine8_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, st, PJ, t_out, n, c);
}

}  // namespace zp

// =====================================================================================================
// Four-step NTT over G ranks.  N = G*M, m = M/G.  With n = n2*G + n1 and k = k1*M + k2:
//     X[k1*M + k2] = sum_{n1} w_G^{n1 k1} * ( w_N^{n1 k2} * sum_{n2} w_M^{n2 k2} x[n2*G + n1] )
// so a rank that holds the CYCLIC slice n1 = r runs one local size-M transform, multiplies by w_N^{r k2},
// exchanges k2-ranges (all-to-all) and finishes with size-G butterflies.  Two more all-to-alls convert the
// caller's contiguous blocks to the cyclic slice and the block-cyclic result back to contiguous blocks; at
// NVLink bandwidth each moves (G-1)/G * 32*M bytes per GPU (0.2 ms at N = 2^25, G = 8).
// =====================================================================================================
namespace zp {

// send[p][i] = in[i*G + p] * (coset ? g^(r*M + i*G + p) : 1)        (block -> cyclic redistribution, packed per peer)
__global__ void ntt4_pack_kernel(const fr_t* __restrict__ in, fr_t* __restrict__ send, size_t M, int lg, size_t base, int coset,
                                 const fr_t* c_lo, const fr_t* c_hi) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    size_t G = (size_t)1 << lg, m = M >> lg;
    fr_t v = load_fr(&in[j]);
    if (coset) {
        size_t n = base + j;
        uint32_t lo = (uint32_t)n & ((1u << NTT_LO_BITS) - 1), hi = (uint32_t)(n >> NTT_LO_BITS);
        if (lo) v = v * load_fr(&c_lo[lo]);
        if (hi) v = v * load_fr(&c_hi[hi]);
    }
    size_t p = j & (G - 1), i = j >> lg;
    store_fr(&send[p * m + i], v);
}
// y[k2] *= w_N^{+-(r * k2)}
__global__ void ntt4_twiddle_kernel(fr_t* __restrict__ y, size_t M, int logn_total, int rank, int inverse, const fr_t* w_lo,
                                    const fr_t* w_hi) {
    size_t k2 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k2 >= M || rank == 0) return;
    uint64_t e = ((uint64_t)rank * k2) & (((uint64_t)1 << logn_total) - 1);
    uint32_t ex = (uint32_t)(e << (NTT_LMAX - logn_total));
    if (!ex) return;
    if (inverse) ex = ((1u << NTT_LMAX) - ex) & ((1u << NTT_LMAX) - 1);
    store_fr(&y[k2], load_fr(&y[k2]) * tw_lookup(w_lo, w_hi, ex));
}
// out[k1][i] = scale * sum_{n1} w_G^{+-(n1 k1)} in[n1][i], G <= 8, radix-2 DIF in registers
__global__ void ntt4_butterfly_kernel(const fr_t* __restrict__ in, fr_t* __restrict__ out, size_t m, int lg, int inverse, int do_scale,
                                      fr_t scale, const fr_t* w_lo, const fr_t* w_hi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int G = 1 << lg;
    fr_t v[8];
    for (int a = 0; a < G; a++) v[a] = load_fr(&in[(size_t)a * m + i]);
    for (int s = 0; s < lg; s++) {
        int half = G >> (s + 1);
        for (int t = 0; t < G / 2; t++) {
            int j = t & (half - 1), blk = t / half;
            int i0 = blk * 2 * half + j, i1 = i0 + half;
            fr_t u = v[i0], w = v[i1];
            v[i0] = u + w;
            fr_t d = u - w;
            if (j) {
                uint32_t ex = ((uint32_t)j << s) << (NTT_LMAX - lg);
                if (inverse) ex = ((1u << NTT_LMAX) - ex) & ((1u << NTT_LMAX) - 1);
                d = d * tw_lookup(w_lo, w_hi, ex);
            }
            v[i1] = d;
        }
    }
    for (int k1 = 0; k1 < G; k1++) {
        int row = lg ? (int)(__brev((uint32_t)k1) >> (32 - lg)) : 0;
        fr_t r = v[row];
        if (do_scale) r = r * scale;
        store_fr(&out[(size_t)k1 * m + i], r);
    }
}
// out[j] = in[j] * g^-(base + j) (coset inverse post-scaling; 1/N already applied)
__global__ void ntt4_coset_out_kernel(fr_t* __restrict__ data, size_t M, size_t base, const fr_t* c_lo, const fr_t* c_hi) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    size_t n = base + j;
    uint32_t lo = (uint32_t)n & ((1u << NTT_LO_BITS) - 1), hi = (uint32_t)(n >> NTT_LO_BITS);
    fr_t v = load_fr(&data[j]);
    if (lo) v = v * load_fr(&c_lo[lo]);
    if (hi) v = v * load_fr(&c_hi[hi]);
    store_fr(&data[j], v);
}

void ntt_sharded_run(const NttTables& T, NttScratch& S, NttKind kind, int logn_total, int rank, int world, const fr_t* in_local,
                     fr_t* out_local, fr_t* tmp_a, fr_t* tmp_b, ntt_alltoall_fn a2a, void* user, cudaStream_t st) {
    int lg = ilog2((size_t)world);
    if (((size_t)1 << lg) != (size_t)world || world > 8) throw std::runtime_error("ntt_sharded_run: world must be 1, 2, 4 or 8");
    if (logn_total > NTT_LMAX || logn_total < 2 * lg) throw std::runtime_error("ntt_sharded_run: unsupported size");
    const bool inverse = (kind == NTT_INV || kind == NTT_COSET_INV);
    if (world == 1) {
        ntt_run(T, S, kind, logn_total, in_local, (size_t)1 << logn_total, out_local, st);
        return;
    }
    const size_t M = (size_t)1 << (logn_total - lg), m = M >> lg;
    const size_t peer_bytes = m * sizeof(fr_t);
    const unsigned gM = (unsigned)((M + 255) / 256), gm = (unsigned)((m + 255) / 256);
    auto exchange = [&](const fr_t* send, fr_t* recv) {
        if (a2a(user, send, recv, peer_bytes) != 0) throw std::runtime_error("ntt_sharded_run: all-to-all failed");
    };
    // 1. block -> cyclic (coset-forward powers g^n folded into the pack)
    ZP_LAUNCH(ntt4_pack_kernel, dim3(gM), dim3(256), 0, st, in_local, tmp_a, M, lg, (size_t)rank * M, kind == NTT_COSET_FWD ? 1 : 0,
              T.g_lo.p, T.g_hi.p);
    exchange(tmp_a, tmp_b);
    // 2. local size-M transform of the cyclic slice (the 1/M of the inverse comes with it)
    ntt_run(T, S, inverse ? NTT_INV : NTT_FWD, logn_total - lg, tmp_b, M, tmp_a, st);
    // 3. twiddle w_N^{r k2}
    ZP_LAUNCH(ntt4_twiddle_kernel, dim3(gM), dim3(256), 0, st, tmp_a, M, logn_total, rank, inverse ? 1 : 0, T.w_lo.p, T.w_hi.p);
    // 4. exchange k2 ranges, 5. size-G butterflies over n1 (times 1/G for the inverse)
    exchange(tmp_a, tmp_b);
    ZP_LAUNCH(ntt4_butterfly_kernel, dim3(gm), dim3(256), 0, st, tmp_b, tmp_a, m, lg, inverse ? 1 : 0, inverse ? 1 : 0, T.ninv[lg],
              T.w_lo.p, T.w_hi.p);
    // 6. block-cyclic -> contiguous blocks
    exchange(tmp_a, out_local);
    if (kind == NTT_COSET_INV)
        ZP_LAUNCH(ntt4_coset_out_kernel, dim3(gM), dim3(256), 0, st, out_local, M, (size_t)rank * M, T.gi_lo.p, T.gi_hi.p);
}

}  // namespace zp
