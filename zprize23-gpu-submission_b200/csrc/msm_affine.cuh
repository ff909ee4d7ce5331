// Batch-affine pre-reduction rounds of the MSM (included by msm.cu).
//
// Before the XYZZ accumulation, the points of every bucket run are added PAIRWISE in affine coordinates:
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1        (2M + 1S + 1 inversion)
// with all inversions of a round shared through one parallel Montgomery batch inversion (3 products per element: the
// leaf level is fused with the denominators on the way up and with the additions on the way down), i.e. ~6 Fq products
// per addition instead of the 10 of an XYZZ mixed addition.  Each round halves the run lengths and
// materialises the partial sums contiguously — all x-coordinates, then all y-coordinates (SoA), so that the next round's
// upward pass streams the x array only — and later rounds and the final accumulation read sequential memory instead of
// gathering from the SRS table.  A pair with x1 == x2 (P + P or P - P; impossible for distinct SRS powers, reachable
// with repeated input points) raises a flag and the whole MSM is redone on the plain XYZZ path, so the result is exact
// for every input.
#pragma once

namespace zp {

static const int BI_CH = 8;  // children per node of the batch-inversion product tree

__global__ void __launch_bounds__(256) ba_pair_count_kernel(const uint32_t* __restrict__ begin, const uint32_t* __restrict__ endp,
                                                            size_t nb, uint32_t* __restrict__ cnt) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint32_t len = endp[b] - begin[b];
    cnt[b] = (len + 1) >> 1;
}

// Slot table: src0[j] = (index of the first point of pair j) << 1 | (the slot is a real pair, not a leftover).
// Eight lanes per bucket run, four runs per warp (coalesced 32-byte stores; no per-slot search): a run has ~52 / 26 / 13 / 7
// pairs in rounds 1 - 4, and one WARP per run spent 0.3 ms per round on two million nearly empty warps.  A run with more than
// 64 pairs (skewed scalar distributions) is written by the whole warp afterwards.
__global__ void __launch_bounds__(256) ba_slots_kernel(const uint32_t* __restrict__ begin, const uint32_t* __restrict__ endp,
                                                       const uint32_t* __restrict__ rs, size_t nb, uint32_t* __restrict__ src0) {
    const size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const uint32_t lane = threadIdx.x & 31, sub = lane & 7;
    uint32_t b0 = 0, len = 0, o = 0;
    if (b < nb) {
        b0 = begin[b];
        len = endp[b] - b0;
        o = rs[b];
    }
    const uint32_t np = (len + 1) >> 1;
    const bool big = np > 64;
    if (!big)
        for (uint32_t t = sub; t < np; t += 8) src0[o + t] = ((b0 + 2 * t) << 1) | (2 * t + 1 < len ? 1u : 0u);
    uint32_t todo = __ballot_sync(0xffffffffu, big && sub == 0);  // every lane gets here: no early return above
    while (todo) {
        const int leader = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t B0 = __shfl_sync(0xffffffffu, b0, leader), L = __shfl_sync(0xffffffffu, len, leader),
                       O = __shfl_sync(0xffffffffu, o, leader);
        const uint32_t NP = (L + 1) >> 1;
        for (uint32_t t = lane; t < NP; t += 32) src0[O + t] = ((B0 + 2 * t) << 1) | (2 * t + 1 < L ? 1u : 0u);
    }
}

// A CTA of BA_T threads covers BA_K * BA_T consecutive slots; thread tl owns slots base + k * BA_T + tl (k < BA_K), so
// every per-slot array is read and written coalesced.  Those BA_K slots form one leaf group of the inversion tree.
static const int BA_K = 16;
static const int BA_T = 256;

// Leaf level of the batch inversion fused with the denominators den = x2 - x1 (1 for a leftover / unused / degenerate
// slot): pre[j] = product of the group's denominators before slot j (k >= 1), up[group] = product of all BA_K.
// The denominators themselves are not stored: the downward kernel reloads both points anyway.
__global__ void __launch_bounds__(BA_T) ba_up0_kernel(const uint32_t* __restrict__ src0, const uint32_t* __restrict__ total_ptr,
                                                      size_t cap, const uint32_t* __restrict__ entries, const PointSrc ps,
                                                      fq_t* __restrict__ pre, fq_t* __restrict__ up,
                                                      uint32_t* __restrict__ flag) {
    const size_t base = (size_t)blockIdx.x * (BA_K * BA_T) + threadIdx.x;
    const size_t total = *total_ptr;
    fq_t acc = fq_t::one();
#pragma unroll 1
    for (int k = 0; k < BA_K; k++) {
        const size_t j = base + (size_t)k * BA_T;
        if (j >= cap) break;
        bool pair = false;
        fq_t d;
        if (j < total) {
            const uint32_t s = src0[j];
            if (s & 1u) {
                const uint32_t i0 = s >> 1;
                const uint32_t e0 = entries ? entries[i0] & 0x7fffffffu : i0, e1 = entries ? entries[i0 + 1] & 0x7fffffffu : i0 + 1;
                d = point_x(ps, e1) - point_x(ps, e0);
                pair = true;
                if (d.is_zero()) {
                    *flag = 1;
                    pair = false;
                }
            }
        }
        if (!pair) d = fq_t::one();
        if (k) store_fq(&pre[j], acc);
        if (k == 0) acc = d;
        else if (pair) acc = acc * d;
    }
    store_fq(&up[(size_t)blockIdx.x * BA_T + threadIdx.x], acc);
}


// (Two variants of this kernel — software prefetch of the next slots' lines with prefetch.global.L2, and 2 / 4 leaf groups
// per thread in lock step — were measured slower, profiles/r02j_msm_up0_variants.log, and removed: the kernel runs at the
// random-line rate of the memory system in round 1.)

// Leaf level downwards fused with the additions: from inv = 1 / (product of the group) recover each 1 / den[j]
// (2 products per slot with the stored prefix products) and emit out[j] = P(i0) + P(i0 + 1), or the leftover point.
__global__ void __launch_bounds__(BA_T, 2) ba_down0_kernel(const uint32_t* __restrict__ src0, const uint32_t* __restrict__ total_ptr,
                                                        size_t cap, const uint32_t* __restrict__ entries, const PointSrc ps,
                                                        const fq_t* __restrict__ pre, const fq_t* __restrict__ up_inv,
                                                        fq_t* __restrict__ out_x, fq_t* __restrict__ out_y) {
    const size_t base = (size_t)blockIdx.x * (BA_K * BA_T) + threadIdx.x;
    const size_t total = *total_ptr;
    if (base >= total) return;  // slots of one thread ascend with k: nothing to emit
    fq_t inv = load_fq(&up_inv[(size_t)blockIdx.x * BA_T + threadIdx.x]);
#pragma unroll 1
    for (int k = BA_K - 1; k >= 0; k--) {
        const size_t j = base + (size_t)k * BA_T;
        if (j >= total) continue;  // unused slots carry den = 1: inv is unchanged
        const uint32_t s = src0[j];
        const uint32_t i0 = s >> 1;
        const uint32_t e0 = entries ? entries[i0] : i0;
        fq_t x1 = point_x(ps, e0 & 0x7fffffffu), y1 = point_y(ps, e0 & 0x7fffffffu);
        if (entries && (e0 >> 31)) y1 = y1.neg();
        if (s & 1u) {
            const uint32_t e1 = entries ? entries[i0 + 1] : i0 + 1;
            fq_t x2 = point_x(ps, e1 & 0x7fffffffu), y2 = point_y(ps, e1 & 0x7fffffffu);
            if (entries && (e1 >> 31)) y2 = y2.neg();
            fq_t d = x2 - x1;
            if (!d.is_zero()) {  // a degenerate pair was given den = 1 (the whole MSM is redone anyway)
                fq_t ik = inv;
                if (k) {
                    ik = inv * load_fq(&pre[j]);
                    inv = inv * d;
                }
                fq_t lam = (y2 - y1) * ik;
                fq_t x3 = lam.sqr() - x1 - x2;
                y1 = lam * (x1 - x3) - y1;
                x1 = x3;
            }
        }
        store_fq(&out_x[j], x1);
        store_fq(&out_y[j], y1);
    }
}

// ---- parallel Montgomery batch inversion over a product tree with BI_CH children per node
__global__ void __launch_bounds__(256) bi_up_kernel(const fq_t* __restrict__ in, size_t n, fq_t* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * BI_CH;
    if (lo >= n) return;
    size_t hi = lo + BI_CH < n ? lo + BI_CH : n;
    fq_t acc = load_fq(&in[lo]);
    for (size_t i = lo + 1; i < hi; i++) acc = acc * load_fq(&in[i]);
    store_fq(&out[t], acc);
}
// vals[lo..hi) <- their inverses, given parent_inv[t] = 1 / prod(vals[lo..hi))
__global__ void __launch_bounds__(128) bi_down_kernel(fq_t* __restrict__ vals, size_t n, const fq_t* __restrict__ parent_inv) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * BI_CH;
    if (lo >= n) return;
    int cnt = (int)((n - lo) < (size_t)BI_CH ? (n - lo) : (size_t)BI_CH);
    fq_t v[BI_CH], pre[BI_CH];
    fq_t acc = fq_t::one();
#pragma unroll
    for (int k = 0; k < BI_CH; k++) {
        if (k < cnt) {
            v[k] = load_fq(&vals[lo + k]);
            pre[k] = acc;
            if (k + 1 < cnt) acc = acc * v[k];
        }
    }
    fq_t inv = load_fq(&parent_inv[t]);
#pragma unroll
    for (int k = BI_CH - 1; k >= 0; k--) {
        if (k < cnt) {
            store_fq(&vals[lo + k], inv * pre[k]);
            if (k > 0) inv = inv * v[k];
        }
    }
}

// In place: d[i] <- 1 / d[i] for i < n (all d[i] != 0).  `levels` holds >= n/7 + 64 scratch elements.  The product tree
// is built on the device down to a level of <= BI_HOST_TOP nodes; that level makes ONE host round trip (the reference
// inverts every element separately, mont_arithmetic.cu:72-78): it comes back into pinned[0 .. TOP), the host inverts it
// with Montgomery's trick (one Fermat inversion + 3 products per node, ~40 us) and the inverses leave from
// pinned[TOP .. 2 TOP) — ONE stream synchronisation per call; the next call's synchronisation orders the host's next write
// after this call's upload has been consumed.  The three device levels this replaces are single-CTA launches that cost
// 25 us going up and 140 us coming down (dependent 255-bit products at ~2 us each), per batch-affine round, at any size.
static const size_t BI_HOST_TOP = 256;
static void fq_batch_inverse(fq_t* d, size_t n, fq_t* levels, fq_t* pinned, cudaStream_t st) {
    if (!n) return;
    std::vector<fq_t*> lv;
    std::vector<size_t> sz;
    lv.push_back(d);
    sz.push_back(n);
    fq_t* next = levels;
    while (sz.back() > BI_HOST_TOP) {
        size_t m = (sz.back() + BI_CH - 1) / BI_CH;
        ZP_LAUNCH(bi_up_kernel, dim3((unsigned)((m + 255) / 256)), dim3(256), 0, st, lv.back(), sz.back(), next);
        lv.push_back(next);
        sz.push_back(m);
        next += m;
    }
    const size_t top = sz.back();
    fq_t* in = pinned;
    fq_t* out = pinned + BI_HOST_TOP;
    ZP_CUDA(cudaMemcpyAsync(in, lv.back(), top * sizeof(fq_t), cudaMemcpyDeviceToHost, st));
    ZP_CUDA(cudaStreamSynchronize(st));
    {
        host::Fq v[BI_HOST_TOP], pre[BI_HOST_TOP];
        host::Fq acc = host::to_host(in[0]);
        v[0] = acc;
        for (size_t i = 1; i < top; i++) {
            v[i] = host::to_host(in[i]);
            pre[i] = acc;          // product of v[0 .. i)
            acc = acc * v[i];
        }
        host::Fq inv = acc.inverse();  // 1 / product of all
        for (size_t i = top - 1; i > 0; i--) {
            out[i] = host::to_dev(inv * pre[i]);
            inv = inv * v[i];
        }
        out[0] = host::to_dev(inv);
    }
    ZP_CUDA(cudaMemcpyAsync(lv.back(), out, top * sizeof(fq_t), cudaMemcpyHostToDevice, st));
    for (size_t l = lv.size() - 1; l-- > 0;) {
        size_t m = sz[l + 1];
        ZP_LAUNCH(bi_down_kernel, dim3((unsigned)((m + 127) / 128)), dim3(128), 0, st, lv[l], sz[l], lv[l + 1]);
    }
}

}  // namespace zp
