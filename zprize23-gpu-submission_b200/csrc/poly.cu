// Element-wise, scan, evaluation and fused quotient kernels — see poly.cuh and DESIGN.md §kernels.
#include "poly.cuh"
#include "gates.cuh"

namespace zp {

static const int EW_BLOCK = 256;
static inline dim3 ew_grid(size_t n) { return dim3((unsigned)((n + EW_BLOCK - 1) / EW_BLOCK)); }

void PolyScratch::init() {
    if (!host_pinned) ZP_CUDA(cudaMallocHost((void**)&host_pinned, 65536));
    if (!flag.p) flag.alloc(4);
}
PolyScratch::~PolyScratch() {
    if (host_pinned) cudaFreeHost(host_pinned);
}

// ------------------------------------------------------------------ lookup helpers
__global__ void compress4_kernel(fr_t* out, const fr_t* a, const fr_t* b, const fr_t* c, const fr_t* d, fr_t zeta, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_fr(&out[i], lc4(load_fr(&a[i]), load_fr(&b[i]), load_fr(&c[i]), load_fr(&d[i]), zeta));
}
void compress4(fr_t* out, const fr_t* a, const fr_t* b, const fr_t* c, const fr_t* d, const fr_t& zeta, size_t n, cudaStream_t st) {
    ZP_LAUNCH(compress4_kernel, ew_grid(n), dim3(EW_BLOCK), 0, st, out, a, b, c, d, zeta, n);
}

__global__ void query_f_kernel(fr_t* out, const fr_t* w0, const fr_t* w1, const fr_t* w2, const fr_t* w3, const fr_t* q_lookup,
                               size_t n_real, const fr_t* t_ev, fr_t zeta, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool on = i < n_real && !load_fr(&q_lookup[i]).is_zero();
    fr_t v;
    if (on) {
        v = lc4(load_fr(&w0[i]), load_fr(&w1[i]), load_fr(&w2[i]), load_fr(&w3[i]), zeta);
    } else {
        v = load_fr(&t_ev[0]);  // (t[0], 0, 0, 0) compressed
    }
    store_fr(&out[i], v);
}
void query_f(fr_t* out, const fr_t* w0, const fr_t* w1, const fr_t* w2, const fr_t* w3, const fr_t* q_lookup, size_t n_real,
             const fr_t* t_ev, const fr_t& zeta, size_t n, cudaStream_t st) {
    ZP_LAUNCH(query_f_kernel, ew_grid(n), dim3(EW_BLOCK), 0, st, out, w0, w1, w2, w3, q_lookup, n_real, t_ev, zeta, n);
}

// ------------------------------------------------------------------ permutation / lookup ratios
struct Ptr4 {
    const fr_t* p[4];
};
__global__ void perm_num_den_kernel(fr_t* num, fr_t* den, Ptr4 w, Ptr4 sigma, fr_t beta, fr_t gamma, int logn, const fr_t* w_lo,
                                    const fr_t* w_hi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> logn) return;
    // root = omega_N^i
    uint32_t ex = (uint32_t)i << (NTT_LMAX - logn);
    uint32_t lo = ex & ((1u << NTT_LO_BITS) - 1), hi = ex >> NTT_LO_BITS;
    fr_t root = load_fr(&w_hi[hi]);
    if (lo) root = root * load_fr(&w_lo[lo]);
    fr_t br = beta * root;
    // K = 1, 7, 13, 17 (permutation/constants.rs:12-22): beta*K*root by additions
    fr_t b2 = br.dbl(), b4 = b2.dbl(), b8 = b4.dbl(), b16 = b8.dbl();
    fr_t k7 = b8 - br, k13 = b8 + b4 + br, k17 = b16 + br;
    fr_t kk[4] = {br, k7, k13, k17};
    fr_t nu = fr_t::one(), de = fr_t::one();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        fr_t wv = load_fr(&w.p[k][i]);
        fr_t wg = wv + gamma;
        fr_t a = wg + kk[k];
        fr_t b = wg + beta * load_fr(&sigma.p[k][i]);
        nu = (k == 0) ? a : nu * a;
        de = (k == 0) ? b : de * b;
    }
    store_fr(&num[i], nu);
    store_fr(&den[i], de);
}
void perm_num_den(fr_t* num, fr_t* den, const fr_t* const w[4], const fr_t* const sigma[4], const fr_t& beta, const fr_t& gamma,
                  int logn, const NttTables& T, cudaStream_t st) {
    Ptr4 pw, ps;
    for (int k = 0; k < 4; k++) {
        pw.p[k] = w[k];
        ps.p[k] = sigma[k];
    }
    size_t n = (size_t)1 << logn;
    ZP_LAUNCH(perm_num_den_kernel, ew_grid(n), dim3(EW_BLOCK), 0, st, num, den, pw, ps, beta, gamma, logn, T.w_lo.p, T.w_hi.p);
}

__global__ void lookup_num_den_kernel(fr_t* num, fr_t* den, const fr_t* f, const fr_t* t, const fr_t* h1, const fr_t* h2, fr_t delta,
                                      fr_t epsilon, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t nx = (i + 1 == n) ? 0 : i + 1;
    fr_t opd = fr_t::one() + delta, eopd = epsilon * opd;
    fr_t fi = load_fr(&f[i]), ti = load_fr(&t[i]), tn = load_fr(&t[nx]);
    fr_t h1i = load_fr(&h1[i]), h1n = load_fr(&h1[nx]), h2i = load_fr(&h2[i]);
    fr_t nu = opd * (epsilon + fi) * (eopd + ti + (delta * tn));
    fr_t de = (eopd + h1i + (h2i * delta)) * (eopd + h2i + (h1n * delta));
    store_fr(&num[i], nu);
    store_fr(&den[i], de);
}
void lookup_num_den(fr_t* num, fr_t* den, const fr_t* f, const fr_t* t, const fr_t* h1, const fr_t* h2, const fr_t& delta,
                    const fr_t& epsilon, size_t n, cudaStream_t st) {
    ZP_LAUNCH(lookup_num_den_kernel, ew_grid(n), dim3(EW_BLOCK), 0, st, num, den, f, t, h1, h2, delta, epsilon, n);
}

// den[i] <- num[i] * den[i]^-1.  RATIO_K elements per thread share one Fermat inversion (Montgomery's trick); a CTA covers
// RATIO_K * blockDim consecutive elements and thread t takes elements base + k * blockDim + t, so every access is coalesced.
// The running prefix products go through `pre` (n elements of scratch) instead of registers: 4 products per element
// + 380 / RATIO_K for the inversion (the previous 8-per-thread register version spent 51 products per element).
static const int RATIO_K = 64;
__global__ void __launch_bounds__(64) ratio_kernel(const fr_t* __restrict__ num, fr_t* __restrict__ den, fr_t* __restrict__ pre, size_t n) {
    const size_t base = (size_t)blockIdx.x * (RATIO_K * blockDim.x) + threadIdx.x;
    if (base >= n) return;
    fr_t acc = fr_t::one();
    int cnt = 0;
#pragma unroll 1
    for (int k = 0; k < RATIO_K; k++) {
        const size_t j = base + (size_t)k * blockDim.x;
        if (j >= n) break;
        store_fr(&pre[j], acc);
        acc = acc * load_fr(&den[j]);
        cnt = k + 1;
    }
    fr_t inv = acc.inverse();
#pragma unroll 1
    for (int k = cnt - 1; k >= 0; k--) {
        const size_t j = base + (size_t)k * blockDim.x;
        fr_t r = inv * load_fr(&pre[j]);
        inv = inv * load_fr(&den[j]);
        store_fr(&den[j], r * load_fr(&num[j]));
    }
}
void ratio_inplace(const fr_t* num, fr_t* den, fr_t* scratch, size_t n, cudaStream_t st) {
    const int threads = 64;
    size_t per_cta = (size_t)RATIO_K * threads;
    ZP_LAUNCH(ratio_kernel, dim3((unsigned)((n + per_cta - 1) / per_cta)), dim3(threads), 0, st, num, den, scratch, n);
}

// ------------------------------------------------------------------ chunked scans
static const int SC_CH = 32;

// part[t] = prod of r over chunk t
__global__ void chunk_product_kernel(const fr_t* r, size_t n, fr_t* part) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * SC_CH;
    if (lo >= n) return;
    size_t hi = lo + SC_CH < n ? lo + SC_CH : n;
    fr_t acc = load_fr(&r[lo]);
    for (size_t i = lo + 1; i < hi; i++) acc = acc * load_fr(&r[i]);
    store_fr(&part[t], acc);
}
// in place exclusive prefix product by one thread (top level, n <= SC_CH)
__global__ void serial_exclusive_product_kernel(const fr_t* r, fr_t* out, size_t n) {
    if (blockIdx.x || threadIdx.x) return;
    fr_t acc = fr_t::one();
    for (size_t i = 0; i < n; i++) {
        fr_t v = load_fr(&r[i]);
        store_fr(&out[i], acc);
        acc = acc * v;
    }
}
// out[i] = carry[t] * prod_{lo <= j < i} r[j]
__global__ void chunk_exclusive_product_kernel(const fr_t* r, size_t n, const fr_t* carry, fr_t* out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * SC_CH;
    if (lo >= n) return;
    size_t hi = lo + SC_CH < n ? lo + SC_CH : n;
    fr_t acc = load_fr(&carry[t]);
    for (size_t i = lo; i < hi; i++) {
        fr_t v = load_fr(&r[i]);
        store_fr(&out[i], acc);
        acc = acc * v;
    }
}
static void exclusive_product_rec(PolyScratch& S, int level, const fr_t* r, fr_t* out, size_t n, cudaStream_t st) {
    if (n <= (size_t)SC_CH) {
        ZP_LAUNCH(serial_exclusive_product_kernel, dim3(1), dim3(32), 0, st, r, out, n);
        return;
    }
    size_t nt = (n + SC_CH - 1) / SC_CH;
    if (S.lv[level].n < nt) S.lv[level].alloc(nt);
    fr_t* part = S.lv[level].p;
    ZP_LAUNCH(chunk_product_kernel, ew_grid(nt), dim3(EW_BLOCK), 0, st, r, n, part);
    exclusive_product_rec(S, level + 1, part, part, nt, st);  // part <- exclusive prefix of chunk products
    ZP_LAUNCH(chunk_exclusive_product_kernel, ew_grid(nt), dim3(EW_BLOCK), 0, st, r, n, part, out);
}
void exclusive_prefix_product(PolyScratch& S, const fr_t* r, fr_t* out, size_t n, cudaStream_t st) {
    exclusive_product_rec(S, 0, r, out, n, st);
}

// Horner chunks: part[t] = sum_{i in chunk t} p[i] z^(i - lo)
__global__ void chunk_horner_kernel(const fr_t* p, size_t n, fr_t z, fr_t* part) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * SC_CH;
    if (lo >= n) return;
    size_t hi = lo + SC_CH < n ? lo + SC_CH : n;
    fr_t acc = load_fr(&p[hi - 1]);
    for (size_t i = hi - 1; i > lo; i--) acc = acc * z + load_fr(&p[i - 1]);
    store_fr(&part[t], acc);
}
// batched variant: polynomial index = blockIdx.y
struct EvalBatch {
    const fr_t* poly[32];
    fr_t point[32];
};
__global__ void chunk_horner_batch_kernel(EvalBatch b, size_t n, fr_t* part, size_t part_stride) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * SC_CH;
    if (lo >= n) return;
    size_t hi = lo + SC_CH < n ? lo + SC_CH : n;
    const fr_t* p = b.poly[blockIdx.y];
    fr_t z = b.point[blockIdx.y];
    fr_t acc = load_fr(&p[hi - 1]);
    for (size_t i = hi - 1; i > lo; i--) acc = acc * z + load_fr(&p[i - 1]);
    store_fr(&part[blockIdx.y * part_stride + t], acc);
}

void evaluate_many(PolyScratch& S, const fr_t* const* polys, const fr_t* points, int count, size_t n, fr_t* results_host,
                   cudaStream_t st) {
    if (count > 32) throw std::runtime_error("evaluate_many: at most 32 evaluations per call");
    S.init();
    EvalBatch b;
    host::Fr zp[32];
    for (int k = 0; k < count; k++) {
        b.poly[k] = polys[k];
        b.point[k] = points[k];
        zp[k] = host::to_host(points[k]);
    }
    size_t cur = n;
    int level = 0;
    // level 0..: reduce by SC_CH per sweep until <= 64 partials per polynomial
    const fr_t* src_base = nullptr;
    size_t src_stride = 0;
    while (cur > 64) {
        size_t nt = (cur + SC_CH - 1) / SC_CH;
        if (S.lv[level].n < nt * count) S.lv[level].alloc(nt * count);
        fr_t* part = S.lv[level].p;
        if (level > 0) {
            for (int k = 0; k < count; k++) b.poly[k] = src_base + k * src_stride;
        }
        ZP_LAUNCH(chunk_horner_batch_kernel, dim3((unsigned)((nt + EW_BLOCK - 1) / EW_BLOCK), count), dim3(EW_BLOCK), 0, st, b, cur,
                  part, nt);
        for (int k = 0; k < count; k++) {
            zp[k] = zp[k].pow_u64(SC_CH);
            b.point[k] = host::to_dev(zp[k]);
        }
        src_base = part;
        src_stride = nt;
        cur = nt;
        level++;
    }
    // finish on the host: count * cur (<= 32 * 64) elements
    std::vector<fr_t> tail((size_t)count * cur);
    if (level == 0) {
        for (int k = 0; k < count; k++)
            ZP_CUDA(cudaMemcpyAsync(tail.data() + (size_t)k * cur, polys[k], cur * sizeof(fr_t), cudaMemcpyDeviceToHost, st));
    } else {
        ZP_CUDA(cudaMemcpyAsync(tail.data(), src_base, (size_t)count * cur * sizeof(fr_t), cudaMemcpyDeviceToHost, st));
    }
    ZP_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < count; k++) {
        host::Fr acc = host::Fr::zero();
        for (size_t i = cur; i-- > 0;) acc = acc * zp[k] + host::to_host(tail[(size_t)k * cur + i]);
        results_host[k] = host::to_dev(acc);
    }
}

// suffix Horner: s_i = p_i + z s_{i+1}.  Level structure as in the prefix product.
// top level: one thread, in place: out[i] = s_i
__global__ void serial_suffix_horner_kernel(const fr_t* p, fr_t* out, size_t n, fr_t z) {
    if (blockIdx.x || threadIdx.x) return;
    fr_t acc = fr_t::zero();
    for (size_t i = n; i-- > 0;) {
        acc = acc * z + load_fr(&p[i]);
        store_fr(&out[i], acc);
    }
}
// given S1[t] = suffix value at chunk granularity (s of chunk t including itself), carry into chunk t is S1[t+1].
// shift = 1: write s_i to q[i-1] (quotient layout), s_0 dropped; shift = 0: out[i] = s_i
__global__ void chunk_suffix_horner_kernel(const fr_t* p, size_t n, fr_t z, const fr_t* s1, size_t nt, fr_t* out, int shift) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * SC_CH;
    if (lo >= n) return;
    size_t hi = lo + SC_CH < n ? lo + SC_CH : n;
    fr_t acc = (t + 1 < nt) ? load_fr(&s1[t + 1]) : fr_t::zero();
    for (size_t i = hi; i-- > lo;) {
        acc = acc * z + load_fr(&p[i]);
        if (shift) {
            if (i > 0) store_fr(&out[i - 1], acc);
        } else {
            store_fr(&out[i], acc);
        }
    }
}
static void suffix_horner_rec(PolyScratch& S, int level, const fr_t* p, fr_t* out, size_t n, const host::Fr& z, int shift,
                              cudaStream_t st) {
    if (n <= (size_t)SC_CH) {
        if (shift) throw std::runtime_error("divide_by_linear: polynomial too short");
        ZP_LAUNCH(serial_suffix_horner_kernel, dim3(1), dim3(32), 0, st, p, out, n, host::to_dev(z));
        return;
    }
    size_t nt = (n + SC_CH - 1) / SC_CH;
    if (S.lv[level].n < nt) S.lv[level].alloc(nt);
    fr_t* part = S.lv[level].p;
    ZP_LAUNCH(chunk_horner_kernel, ew_grid(nt), dim3(EW_BLOCK), 0, st, p, n, host::to_dev(z), part);
    host::Fr zc = z.pow_u64(SC_CH);
    suffix_horner_rec(S, level + 1, part, part, nt, zc, 0, st);  // part[t] <- suffix value at chunk t
    ZP_LAUNCH(chunk_suffix_horner_kernel, ew_grid(nt), dim3(EW_BLOCK), 0, st, p, n, host::to_dev(z), part, nt, out, shift);
}
void divide_by_linear(PolyScratch& S, const fr_t* p, size_t n, const fr_t& z, fr_t* q, cudaStream_t st) {
    if (n <= (size_t)SC_CH) {
        // tiny polynomials: go through a padded copy so the chunked path applies
        throw std::runtime_error("divide_by_linear: n must exceed 32");
    }
    suffix_horner_rec(S, 0, p, q, n, host::to_host(z), 1, st);
    ZP_CUDA(cudaMemsetAsync(q + (n - 1), 0, sizeof(fr_t), st));
}

// ------------------------------------------------------------------ linear combination
struct LinBatch {
    const fr_t* poly[40];
    fr_t s[40];
    int count;
};
__global__ void lincomb_kernel(fr_t* out, LinBatch b, size_t n, int accumulate) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t acc = accumulate ? load_fr(&out[i]) : fr_t::zero();
    for (int k = 0; k < b.count; k++) acc = acc + load_fr(&b.poly[k][i]) * b.s[k];
    store_fr(&out[i], acc);
}
void lincomb(fr_t* out, const fr_t* const* polys, const fr_t* scalars, int count, size_t n, cudaStream_t st) {
    int done = 0;
    bool first = true;
    if (count == 0) {
        ZP_CUDA(cudaMemsetAsync(out, 0, n * sizeof(fr_t), st));
        return;
    }
    while (done < count) {
        LinBatch b;
        b.count = count - done < 40 ? count - done : 40;
        for (int k = 0; k < b.count; k++) {
            b.poly[k] = polys[done + k];
            b.s[k] = scalars[done + k];
        }
        ZP_LAUNCH(lincomb_kernel, ew_grid(n), dim3(EW_BLOCK), 0, st, out, b, n, first ? 0 : 1);
        first = false;
        done += b.count;
    }
}

// ------------------------------------------------------------------ flags
__global__ void any_nonzero_kernel(const fr_t* a, size_t n, uint32_t* flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!load_fr(&a[i]).is_zero()) *flag = 1;
}
__global__ void any_differs_from_first_kernel(const fr_t* a, size_t n, uint32_t* flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (load_fr(&a[i]) != load_fr(&a[0])) *flag = 1;
}
static bool read_flag(PolyScratch& S, cudaStream_t st) {
    uint32_t h = 0;
    ZP_CUDA(cudaMemcpyAsync(&h, S.flag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    ZP_CUDA(cudaStreamSynchronize(st));
    return h != 0;
}
bool all_zero(PolyScratch& S, const fr_t* a, size_t n, cudaStream_t st) {
    S.init();
    ZP_CUDA(cudaMemsetAsync(S.flag.p, 0, sizeof(uint32_t), st));
    ZP_LAUNCH(any_nonzero_kernel, ew_grid(n), dim3(EW_BLOCK), 0, st, a, n, S.flag.p);
    return !read_flag(S, st);
}
bool all_equal_to_first(PolyScratch& S, const fr_t* a, size_t n, cudaStream_t st) {
    S.init();
    ZP_CUDA(cudaMemsetAsync(S.flag.p, 0, sizeof(uint32_t), st));
    ZP_LAUNCH(any_differs_from_first_kernel, ew_grid(n), dim3(EW_BLOCK), 0, st, a, n, S.flag.p);
    return !read_flag(S, st);
}
__global__ void fill_kernel(fr_t* out, fr_t v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_fr(&out[i], v);
}
void fill(fr_t* out, const fr_t& v, size_t n, cudaStream_t st) { ZP_LAUNCH(fill_kernel, ew_grid(n), dim3(EW_BLOCK), 0, st, out, v, n); }

// ------------------------------------------------------------------ plookup combine_split on the device
static const uint32_t HS_EMPTY = 0xffffffffu;
ZP_D uint32_t fr_hash(const fr_t& v) {
    uint32_t h = v.l[0] * 0x9e3779b1u;
    h ^= (v.l[1] + 0x7f4a7c15u) * 0x85ebca6bu;
    h ^= (v.l[3] ^ (h >> 15)) * 0xc2b2ae35u;
    h ^= v.l[6] * 0x27d4eb2fu;
    return h ^ (h >> 16);
}
__global__ void cs_zero_counts_kernel(uint32_t* __restrict__ table, size_t slots) {
    size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < slots) table[2 * s + 1] = 0;
}
// table[2s] = smallest index in t holding this value, table[2s+1] = occurrences in t and f
__global__ void cs_insert_t_kernel(const fr_t* __restrict__ t, size_t n, uint32_t* __restrict__ table, uint32_t mask) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t v = load_fr(&t[i]);
    uint32_t s = fr_hash(v) & mask;
    while (true) {
        uint32_t old = atomicCAS(&table[2 * s], HS_EMPTY, (uint32_t)i);
        if (old == HS_EMPTY || load_fr(&t[old]) == v) {
            if (old != HS_EMPTY) atomicMin(&table[2 * s], (uint32_t)i);
            atomicAdd(&table[2 * s + 1], 1u);
            return;
        }
        s = (s + 1) & mask;
    }
}
__global__ void cs_count_f_kernel(const fr_t* __restrict__ f, size_t n, const fr_t* __restrict__ t, uint32_t* __restrict__ table,
                                  uint32_t mask, uint32_t* __restrict__ err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t v = load_fr(&f[i]);
    uint32_t s = fr_hash(v) & mask;
    while (true) {
        uint32_t idx = table[2 * s];
        if (idx == HS_EMPTY) {
            *err = 1;  // Error::ElementNotIndexed
            return;
        }
        if (load_fr(&t[idx]) == v) {
            atomicAdd(&table[2 * s + 1], 1u);
            return;
        }
        s = (s + 1) & mask;
    }
}
// cnt[first index] = bucket size
__global__ void cs_emit_counts_kernel(const uint32_t* __restrict__ table, size_t slots, uint32_t* __restrict__ cnt) {
    size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= slots) return;
    uint32_t idx = table[2 * s];
    if (idx != HS_EMPTY) cnt[idx] = table[2 * s + 1];
}
// position p of the sorted union belongs to the bucket i with offs[i] <= p < offs[i+1]; even p -> h1, odd p -> h2
__global__ void cs_expand_kernel(const fr_t* __restrict__ t, const uint32_t* __restrict__ offs, size_t n, size_t total,
                                 fr_t* __restrict__ h1, fr_t* __restrict__ h2) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    size_t lo = 0, hi = n;  // largest i < n with offs[i] <= p
    while (hi - lo > 1) {
        size_t mid = (lo + hi) >> 1;
        if (offs[mid] <= (uint32_t)p) lo = mid; else hi = mid;
    }
    fr_t v = load_fr(&t[lo]);
    store_fr((p & 1) ? &h2[p >> 1] : &h1[p >> 1], v);
}
bool combine_split(CombineSplitScratch& S, const fr_t* t, const fr_t* f, size_t n, fr_t* h1, fr_t* h2, cudaStream_t st) {
    return combine_split(S, t, n, f, n, h1, h2, st);
}
// general lengths: t has n elements, f has nf; h1 receives ceil((n + nf) / 2) elements, h2 floor((n + nf) / 2)
bool combine_split(CombineSplitScratch& S, const fr_t* t, size_t n, const fr_t* f, size_t nf, fr_t* h1, fr_t* h2, cudaStream_t st) {
    size_t slots = 1;
    while (slots < 2 * n) slots <<= 1;
    if (S.table.n < 2 * slots) S.table.alloc(2 * slots);
    if (S.cnt.n < n) S.cnt.alloc(n);
    if (S.offs.n < n + 1) S.offs.alloc(n + 1);
    if (S.tile_sum.n < n / 2048 + 2) S.tile_sum.alloc(n / 2048 + 2);
    if (!S.flag.p) S.flag.alloc(1);
    // (idx, cnt) pairs: idx = EMPTY (all ones), cnt = 0
    ZP_CUDA(cudaMemsetAsync(S.table.p, 0xff, 2 * slots * sizeof(uint32_t), st));
    ZP_CUDA(cudaMemsetAsync(S.cnt.p, 0, n * sizeof(uint32_t), st));
    ZP_CUDA(cudaMemsetAsync(S.flag.p, 0, sizeof(uint32_t), st));
    ZP_LAUNCH(cs_zero_counts_kernel, ew_grid(slots), dim3(EW_BLOCK), 0, st, S.table.p, slots);
    ZP_LAUNCH(cs_insert_t_kernel, ew_grid(n), dim3(EW_BLOCK), 0, st, t, n, S.table.p, (uint32_t)(slots - 1));
    if (nf) ZP_LAUNCH(cs_count_f_kernel, ew_grid(nf), dim3(EW_BLOCK), 0, st, f, nf, t, S.table.p, (uint32_t)(slots - 1), S.flag.p);
    ZP_LAUNCH(cs_emit_counts_kernel, ew_grid(slots), dim3(EW_BLOCK), 0, st, S.table.p, slots, S.cnt.p);
    u32_exclusive_scan(S.cnt.p, S.offs.p, n, S.tile_sum.p, st);
    uint32_t err = 0;
    ZP_CUDA(cudaMemcpyAsync(&err, S.flag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    ZP_CUDA(cudaStreamSynchronize(st));
    if (err) return false;
    ZP_LAUNCH(cs_expand_kernel, ew_grid(n + nf), dim3(EW_BLOCK), 0, st, t, S.offs.p, n, n + nf, h1, h2);
    return true;
}

// ------------------------------------------------------------------ fused quotient pass
// One thread per point of the 8N coset.  Reads every stream once (the "+8" rotations hit lines that
// are already on their way through L2), evaluates gate + permutation + lookup constraints and multiplies
// by the 8-periodic inverse of Z_H:  quotient_poly.rs:195-201.
template <bool CUSTOM, bool LOOKUP>
__global__ void __launch_bounds__(128) quotient_kernel(QuotientArgs a) {
    const size_t n8 = (size_t)1 << (a.logn + 3);
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.i_count) return;
    t += a.i_begin;
    // i: natural index on the 8N domain (prover-key streams, l1, x, Z_H); wi / nx: index of this point and of the "next"
    // point (omega * x) in the wire-like arrays
    // ki: index into the prover-key streams and l1 (natural 8N order, or this coset's compact copy); kmask / kshift: the
    // public-input rotation in that indexing
    size_t i = t, wi = t, nx = (t + 8) & (n8 - 1), ki = t, kmask = n8 - 1, kshift = a.pi_shift;
    if (a.coset_j >= 0) {
        const size_t nn = n8 >> 3, c = t & (((size_t)1 << a.coset_lc) - 1), tt = t >> a.coset_lc;
        i = 8 * tt + (size_t)a.coset_j + c;
        wi = c * nn + tt;
        nx = c * nn + ((tt + 1) & (nn - 1));
        ki = i;
        if (a.compact_key) {
            ki = tt;
            kmask = nn - 1;
            kshift = a.pi_shift >> 3;
        }
    }
    const fr_t one = fr_t::one();
    GateVals<fr_t> g;
    g.a = load_fr(&a.w[0][wi]);
    g.b = load_fr(&a.w[1][wi]);
    g.c = load_fr(&a.w[2][wi]);
    g.d = load_fr(&a.w[3][wi]);

    // ---- arithmetic widget (arithmetic.rs:61-79)
    fr_t arith = a.sel[5] ? load_fr(&a.sel[5][ki]) : fr_t::zero();  // q_c (dropped when identically zero, like every selector)
    g.q_c = arith;
    if (a.sel[0]) arith = arith + g.a * g.b * load_fr(&a.sel[0][ki]);
    g.q_l = a.sel[1] ? load_fr(&a.sel[1][ki]) : fr_t::zero();
    g.q_r = a.sel[2] ? load_fr(&a.sel[2][ki]) : fr_t::zero();
    if (a.sel[1]) arith = arith + g.a * g.q_l;
    if (a.sel[2]) arith = arith + g.b * g.q_r;
    if (a.sel[3]) arith = arith + g.c * load_fr(&a.sel[3][ki]);
    if (a.sel[4]) arith = arith + g.d * load_fr(&a.sel[4][ki]);
    if (a.sel[6]) arith = arith + g.a.pow5() * load_fr(&a.sel[6][ki]);
    if (a.sel[7]) arith = arith + g.b.pow5() * load_fr(&a.sel[7][ki]);
    if (a.sel[8]) arith = arith + g.d.pow5() * load_fr(&a.sel[8][ki]);
    fr_t total = a.sel[9] ? arith * load_fr(&a.sel[9][ki]) : fr_t::zero();
    if (a.pi_count) total = total + a.pi_val * load_fr(&a.l1[(ki - kshift) & kmask]);

    if (CUSTOM) {
        g.a_next = load_fr(&a.w[0][nx]);
        g.b_next = load_fr(&a.w[1][nx]);
        g.d_next = load_fr(&a.w[3][nx]);
        if (a.sel[10]) total = total + load_fr(&a.sel[10][ki]) * range_constraints(a.range_sep, g);
        if (a.sel[11]) total = total + load_fr(&a.sel[11][ki]) * logic_constraints(a.logic_sep, g);
        if (a.sel[12]) total = total + load_fr(&a.sel[12][ki]) * fbsm_constraints(a.fixed_sep, g, a.coeff_d);
        if (a.sel[13]) total = total + load_fr(&a.sel[13][ki]) * curve_add_constraints(a.var_sep, g, a.coeff_d);
    }

    // ---- permutation widget (permutation.rs:62-153); x = g * omega_8N^i  (= linear_evaluations[i])
    {
        uint32_t ex = (uint32_t)i << (NTT_LMAX - (a.logn + 3));
        uint32_t lo = ex & ((1u << NTT_LO_BITS) - 1), hi = ex >> NTT_LO_BITS;
        fr_t bx = a.beta_g * load_fr(&a.w_hi[hi]);  // beta * x, beta * g folded on the host
        if (lo) bx = bx * load_fr(&a.w_lo[lo]);
        fr_t b2 = bx.dbl(), b4 = b2.dbl(), b8 = b4.dbl(), b16 = b8.dbl();
        fr_t ag = g.a + a.gamma, bg = g.b + a.gamma, cg = g.c + a.gamma, dg = g.d + a.gamma;
        fr_t zi = load_fr(&a.z[wi]), zn = load_fr(&a.z[nx]);
        fr_t id = (ag + bx) * (bg + (b8 - bx)) * (cg + (b8 + b4 + bx)) * (dg + (b16 + bx)) * zi;
        fr_t cp = (ag + a.beta * load_fr(&a.sigma[0][ki])) * (bg + a.beta * load_fr(&a.sigma[1][ki])) *
                  (cg + a.beta * load_fr(&a.sigma[2][ki])) * (dg + a.beta * load_fr(&a.sigma[3][ki])) * zn;
        fr_t l1a = load_fr(&a.l1[ki]) * a.alpha_sq;
        total = total + (id - cp) * a.alpha + (zi - one) * l1a;
    }

    // ---- lookup widget (lookup.rs:98-152)
    if (LOOKUP) {
        fr_t lsep_sq = a.lookup_sep.sqr(), lsep_cu = lsep_sq * a.lookup_sep;
        fr_t opd = a.delta + one, eopd = a.epsilon * opd;
        fr_t fi = load_fr(&a.f[wi]);
        fr_t ti = load_fr(&a.table[wi]), tn = load_fr(&a.table[nx]);
        fr_t h1i = load_fr(&a.h1[wi]), h1n = load_fr(&a.h1[nx]), h2i = load_fr(&a.h2[wi]);
        fr_t z2i = load_fr(&a.z2[wi]), z2n = load_fr(&a.z2[nx]);
        fr_t la = fr_t::zero();
        if (a.sel[14]) la = load_fr(&a.sel[14][ki]) * (lc4(g.a, g.b, g.c, g.d, a.zeta) - fi) * a.lookup_sep;
        fr_t lb = z2i * opd * (a.epsilon + fi) * (eopd + ti + a.delta * tn) * lsep_sq;
        fr_t lcv = z2n * (eopd + h1i + a.delta * h2i) * (eopd + h2i + a.delta * h1n) * lsep_sq;
        fr_t ld = (z2i - one) * load_fr(&a.l1[ki]) * lsep_cu;
        total = total + la + lb - lcv + ld;
    }
    store_fr(&a.out[wi], total * a.vh_inv[i & 7]);
}

void quotient_evals(const QuotientArgs& a, cudaStream_t st) {
    size_t n8 = (size_t)1 << (a.logn + 3);
    bool custom = a.sel[10] || a.sel[11] || a.sel[12] || a.sel[13];
    bool lookup = a.z2 != nullptr;
    (void)n8;
    dim3 grid((unsigned)((a.i_count + 127) / 128)), block(128);
    if (custom && lookup) {
        auto k = quotient_kernel<true, true>;
        ZP_LAUNCH(k, grid, block, 0, st, a);
    } else if (custom) {
        auto k = quotient_kernel<true, false>;
        ZP_LAUNCH(k, grid, block, 0, st, a);
    } else if (lookup) {
        auto k = quotient_kernel<false, true>;
        ZP_LAUNCH(k, grid, block, 0, st, a);
    } else {
        auto k = quotient_kernel<false, false>;
        ZP_LAUNCH(k, grid, block, 0, st, a);
    }
}

}  // namespace zp
