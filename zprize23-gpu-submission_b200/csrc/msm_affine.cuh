// Batch-affine pre-reduction rounds of the MSM (included by msm.cu).
//
// Before the XYZZ accumulation, the points of every bucket run are added PAIRWISE in affine coordinates:
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1        (2M + 1S + 1 inversion)
// with all inversions of a round shared through one parallel Montgomery batch inversion (3.4 products per element), i.e.
// ~6.5 Fq products per addition instead of the 10 of an XYZZ mixed addition.  Each round halves the run lengths and
// materialises the partial sums contiguously, so later rounds and the final accumulation read sequential memory instead
// of gathering from the SRS table.  A pair with x1 == x2 (P + P or P - P; impossible for distinct SRS powers, reachable
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

// bucket b with rs[b] <= j < rs[b+1]
ZP_D size_t ba_find_bucket(const uint32_t* __restrict__ rs, size_t nb, uint32_t j) {
    size_t lo = 0, hi = nb;
    while (hi - lo > 1) {
        size_t mid = (lo + hi) >> 1;
        if (rs[mid] <= j) lo = mid; else hi = mid;
    }
    return lo;
}

// den[j] = x2 - x1 of the j-th output slot (1 for an unpaired leftover or an unused slot); remembers the source index
__global__ void __launch_bounds__(256) ba_pair_denoms_kernel(const uint32_t* __restrict__ begin, const uint32_t* __restrict__ endp,
                                                             const uint32_t* __restrict__ rs, size_t nb, size_t cap,
                                                             const uint32_t* __restrict__ entries, const affine_t* __restrict__ pts,
                                                             fq_t* __restrict__ den, uint32_t* __restrict__ src0,
                                                             uint32_t* __restrict__ flag) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cap) return;
    fq_t d = fq_t::one();
    uint32_t s = 0xffffffffu;  // unused slot
    if (j < rs[nb]) {
        size_t b = ba_find_bucket(rs, nb, (uint32_t)j);
        uint32_t t = (uint32_t)j - rs[b], len = endp[b] - begin[b];
        uint32_t i0 = begin[b] + 2 * t;
        s = i0 << 1;  // bit 0: this slot is a real pair
        if (2 * t + 1 < len) {
            s |= 1u;
            uint32_t e0 = entries ? entries[i0] & 0x7fffffffu : i0, e1 = entries ? entries[i0 + 1] & 0x7fffffffu : i0 + 1;
            d = load_fq(&pts[e1].x) - load_fq(&pts[e0].x);
            if (d.is_zero()) {
                *flag = 1;
                d = fq_t::one();
            }
        }
    }
    store_fq(&den[j], d);
    src0[j] = s;
}

// out[j] = P(i0) + P(i0 + 1) using inv[j] = 1 / (x2 - x1), or a copy of the leftover point
__global__ void __launch_bounds__(256) ba_pair_sums_kernel(const uint32_t* __restrict__ src0, size_t cap,
                                                           const uint32_t* __restrict__ entries, const affine_t* __restrict__ pts,
                                                           const fq_t* __restrict__ inv, affine_t* __restrict__ out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cap) return;
    uint32_t s = src0[j];
    if (s == 0xffffffffu) return;
    uint32_t i0 = s >> 1;
    uint32_t e0 = entries ? entries[i0] : i0;
    fq_t x1 = load_fq(&pts[e0 & 0x7fffffffu].x), y1 = load_fq(&pts[e0 & 0x7fffffffu].y);
    if (entries && (e0 >> 31)) y1 = y1.neg();
    if (s & 1u) {
        uint32_t e1 = entries ? entries[i0 + 1] : i0 + 1;
        fq_t x2 = load_fq(&pts[e1 & 0x7fffffffu].x), y2 = load_fq(&pts[e1 & 0x7fffffffu].y);
        if (entries && (e1 >> 31)) y2 = y2.neg();
        fq_t lam = (y2 - y1) * load_fq(&inv[j]);
        fq_t x3 = lam.sqr() - x1 - x2;
        fq_t y3 = lam * (x1 - x3) - y1;
        x1 = x3;
        y1 = y3;
    }
    store_fq(&out[j].x, x1);
    store_fq(&out[j].y, y1);
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

// In place: d[i] <- 1 / d[i] for i < n (all d[i] != 0).  `levels` holds >= n/7 + 64 scratch elements.  One host round
// trip inverts the single root product (the reference inverts every element separately, mont_arithmetic.cu:72-78).
static void fq_batch_inverse(fq_t* d, size_t n, fq_t* levels, cudaStream_t st) {
    std::vector<fq_t*> lv;
    std::vector<size_t> sz;
    lv.push_back(d);
    sz.push_back(n);
    fq_t* next = levels;
    while (sz.back() > 1) {
        size_t m = (sz.back() + BI_CH - 1) / BI_CH;
        ZP_LAUNCH(bi_up_kernel, dim3((unsigned)((m + 255) / 256)), dim3(256), 0, st, lv.back(), sz.back(), next);
        lv.push_back(next);
        sz.push_back(m);
        next += m;
    }
    fq_t root;
    ZP_CUDA(cudaMemcpyAsync(&root, lv.back(), sizeof(fq_t), cudaMemcpyDeviceToHost, st));
    ZP_CUDA(cudaStreamSynchronize(st));
    root = host::to_dev(host::to_host(root).inverse());
    ZP_CUDA(cudaMemcpyAsync(lv.back(), &root, sizeof(fq_t), cudaMemcpyHostToDevice, st));
    ZP_CUDA(cudaStreamSynchronize(st));  // `root` lives on this stack frame
    for (size_t l = lv.size() - 1; l-- > 0;) {
        size_t m = sz[l + 1];
        ZP_LAUNCH(bi_down_kernel, dim3((unsigned)((m + 127) / 128)), dim3(128), 0, st, lv[l], sz[l], lv[l + 1]);
    }
}

}  // namespace zp
