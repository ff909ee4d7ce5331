// Pippenger MSM over BLS12-381 G1 for sm_100a — see msm.cuh / DESIGN.md §MSM.
//
// Pipeline (all on one stream, no host round trip until the 16 window sums come back):
//   1. digits    : Montgomery scalar -> canonical -> signed c-bit digits, per-(window,bucket) histogram
//   2. scan      : exclusive scan of the histogram (bucket start offsets), one CTA
//   3. scatter   : counting-sort the point indices into (window,bucket) runs (atomic cursor per bucket)
//   4. accumulate: one thread per work segment (<= 2x mean bucket load) of a bucket, XYZZ mixed additions
//                  (the IMAD-bound hot loop); long buckets are split so no digit distribution serialises
//   5. reduce    : per window and bucket group, running-sum reduction  sum_b (b+1) * B_b  + tree in smem
//   host         : fold 8 partials per window, Horner over windows (256 doublings), to affine
// Order inside a bucket is not deterministic (atomics) but the group sum is exact, so the affine
// result is bit-identical run to run.
#include "msm.cuh"

namespace zp {

MsmConfig msm_config_for(size_t n, int c_override) {
    MsmConfig cfg;
    int lg = ilog2(n < 2 ? 2 : n);
    int c = lg - 5;
    if (c < 5) c = 5;
    if (c > 16) c = 16;
    if (c_override) c = c_override;
    cfg.c = c;
    cfg.nwin = (256 + c - 1) / c;
    cfg.nbuckets = 1 << (c - 1);
    return cfg;
}

void MsmWorkspace::reserve(size_t n, const MsmConfig& cfg) {
    size_t wn = (size_t)cfg.nwin * n, wb = (size_t)cfg.nwin * cfg.nbuckets;
    if (digits.n < wn) digits.alloc(wn);
    if (sorted.n < wn) sorted.alloc(wn);
    if (start.n < wb + 1) start.alloc(wb + 1);
    if (cursor.n < wb) cursor.alloc(wb);
    if (seg_start.n < wb + 1) seg_start.alloc(wb + 1);
    if (seg_cnt.n < wb) seg_cnt.alloc(wb);
    size_t mean = (n + cfg.nbuckets - 1) / cfg.nbuckets;
    seg = 2 * mean < 32 ? 32 : 2 * mean;
    max_segs = wn / seg + wb + 1;
    if (segs.n < max_segs) segs.alloc(max_segs);
    size_t np = (size_t)cfg.nwin * MSM_REDUCE_GROUPS;
    if (partial.n < np) partial.alloc(np);
    if (partial_host.size() < np) partial_host.resize(np);
}

__global__ void __launch_bounds__(256) msm_digits_kernel(const fr_t* __restrict__ scalars, size_t n, int c, int nwin, int nbuckets,
                                                         uint32_t* __restrict__ digits, uint32_t* __restrict__ hist) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t s = load_fr(&scalars[i]).from_mont();
    uint32_t carry = 0;
    const uint32_t mask = (1u << c) - 1, half = 1u << (c - 1);
    for (int w = 0; w < nwin; w++) {
        int bit = w * c;
        int li = bit >> 5, off = bit & 31;
        uint32_t raw = 0;
        if (li < 8) {
            raw = s.l[li] >> off;
            if (off + c > 32 && li + 1 < 8) raw |= s.l[li + 1] << (32 - off);
            raw &= mask;
        }
        uint32_t d = raw + carry;
        uint32_t neg = 0;
        if (d > half) {
            d = (1u << c) - d;
            neg = 1;
            carry = 1;
        } else {
            carry = 0;
        }
        digits[(size_t)w * n + i] = d | (neg << 31);
        if (d) atomicAdd(&hist[(size_t)w * nbuckets + d - 1], 1u);
    }
}

// start[0..m] = exclusive scan of cnt[0..m); cnt[i] <- start[i] (becomes the scatter cursor)
__global__ void __launch_bounds__(1024) msm_scan_kernel(uint32_t* __restrict__ cnt, uint32_t* __restrict__ start, size_t m) {
    __shared__ uint32_t part[1024];
    const int t = threadIdx.x;
    size_t chunk = (m + 1023) / 1024;
    size_t lo = (size_t)t * chunk, hi = lo + chunk < m ? lo + chunk : m;
    uint32_t s = 0;
    for (size_t i = lo; i < hi; i++) s += cnt[i];
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        uint32_t v = (t >= d) ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - s;  // exclusive prefix of this chunk
    for (size_t i = lo; i < hi; i++) {
        uint32_t cval = cnt[i];
        start[i] = run;
        cnt[i] = run;
        run += cval;
    }
    if (t == 1023) start[m] = part[1023];
}

__global__ void __launch_bounds__(256) msm_scatter_kernel(const uint32_t* __restrict__ digits, size_t n, int nwin, int nbuckets,
                                                          uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int w = blockIdx.y;
    if (i >= n) return;
    uint32_t dg = digits[(size_t)w * n + i];
    uint32_t d = dg & 0x7fffffffu;
    if (!d) return;
    uint32_t pos = atomicAdd(&cursor[(size_t)w * nbuckets + d - 1], 1u);
    sorted[pos] = (uint32_t)i | (dg & 0x80000000u);
}

ZP_D affine_t load_affine(const affine_t* p) {
    affine_t r;
    r.x = load_fq(&p->x);
    r.y = load_fq(&p->y);
    return r;
}
ZP_D void store_xyzz(xyzz_t* p, const xyzz_t& v) {
    store_fq(&p->X, v.X);
    store_fq(&p->Y, v.Y);
    store_fq(&p->ZZ, v.ZZ);
    store_fq(&p->ZZZ, v.ZZZ);
}
ZP_D xyzz_t load_xyzz(const xyzz_t* p) {
    xyzz_t r;
    r.X = load_fq(&p->X);
    r.Y = load_fq(&p->Y);
    r.ZZ = load_fq(&p->ZZ);
    r.ZZZ = load_fq(&p->ZZZ);
    return r;
}

// Work segments: a bucket with cnt points is split into ceil(cnt / seg) segments so that one thread never
// walks more than `seg` points, whatever the digit distribution (short top window, repeated scalars …).
__global__ void __launch_bounds__(256) msm_segcount_kernel(const uint32_t* __restrict__ start, const uint32_t* __restrict__ endp,
                                                           size_t nb, uint32_t seg, uint32_t* __restrict__ seg_cnt) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb) return;
    uint32_t c = endp[t] - start[t];
    seg_cnt[t] = (c + seg - 1) / seg;
}

// One thread per work segment: XYZZ mixed additions over <= seg points of one bucket.
__global__ void __launch_bounds__(128) msm_accumulate_kernel(const affine_t* __restrict__ points, const uint32_t* __restrict__ sorted,
                                                             const uint32_t* __restrict__ start, const uint32_t* __restrict__ endp,
                                                             const uint32_t* __restrict__ seg_start, size_t nb, uint32_t seg,
                                                             xyzz_t* __restrict__ segs) {
    size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= seg_start[nb]) return;
    // bucket b with seg_start[b] <= s < seg_start[b+1]
    size_t lo_b = 0, hi_b = nb;
    while (hi_b - lo_b > 1) {
        size_t mid = (lo_b + hi_b) >> 1;
        if (seg_start[mid] <= (uint32_t)s) lo_b = mid; else hi_b = mid;
    }
    uint32_t k0 = start[lo_b] + ((uint32_t)s - seg_start[lo_b]) * seg;
    uint32_t k1 = k0 + seg < endp[lo_b] ? k0 + seg : endp[lo_b];
    xyzz_t acc = xyzz_t::infinity();
    for (uint32_t k = k0; k < k1; k++) {
        uint32_t e = sorted[k];
        affine_t p = load_affine(&points[e & 0x7fffffffu]);
        if (e >> 31) p.y = p.y.neg();
        acc.add_affine(p.x, p.y);
    }
    store_xyzz(&segs[s], acc);
}

ZP_D xyzz_t load_bucket(const xyzz_t* __restrict__ segs, const uint32_t* __restrict__ seg_start, size_t b) {
    uint32_t s0 = seg_start[b], s1 = seg_start[b + 1];
    if (s0 == s1) return xyzz_t::infinity();
    xyzz_t acc = load_xyzz(&segs[s0]);
    for (uint32_t s = s0 + 1; s < s1; s++) acc.add(load_xyzz(&segs[s]));
    return acc;
}

// One CTA per (window, bucket group).  partial[w * G + g] = sum_{b in group} (b + 1) * bucket[w][b]
__global__ void __launch_bounds__(128) msm_reduce_kernel(const xyzz_t* __restrict__ segs, const uint32_t* __restrict__ seg_start,
                                                         int nbuckets, int groups, xyzz_t* __restrict__ partial) {
    ZP_DYN_SMEM(xyzz_t, sm);
    const int w = blockIdx.x / groups, g = blockIdx.x % groups;
    const int bg = nbuckets / groups;          // buckets per group
    const int T = blockDim.x;
    const int L = bg / T;                      // buckets per thread (host guarantees divisibility, L >= 1)
    const int j0 = g * bg + threadIdx.x * L;   // first bucket of this thread (weight j0 + 1)
    const size_t B0 = (size_t)w * nbuckets;
    xyzz_t run = xyzz_t::infinity(), sum = xyzz_t::infinity();
    for (int j = j0 + L - 1; j >= j0; j--) {
        xyzz_t b = load_bucket(segs, seg_start, B0 + j);
        run.add(b);
        sum.add(run);
    }
    // sum = sum_j (j - j0 + 1) B_j ;  add j0 * run
    if (j0) {
        xyzz_t acc = xyzz_t::infinity();
        for (int bit = 31 - __clz((uint32_t)j0); bit >= 0; bit--) {
            acc.dbl_inplace();
            if ((j0 >> bit) & 1) acc.add(run);
        }
        sum.add(acc);
    }
    sm[threadIdx.x] = sum;
    __syncthreads();
    for (int d = T >> 1; d >= 1; d >>= 1) {
        if ((int)threadIdx.x < d) {
            xyzz_t a = sm[threadIdx.x];
            a.add(sm[threadIdx.x + d]);
            sm[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) store_xyzz(&partial[blockIdx.x], sm[0]);
}

void msm_launch(MsmWorkspace& ws, const MsmConfig& cfg, const affine_t* points, const fr_t* scalars, size_t n, cudaStream_t st) {
    ws.reserve(n, cfg);
    const size_t wb = (size_t)cfg.nwin * cfg.nbuckets;
    auto mark = [&](int k) {
        if (!ws.timing) return;
        if (!ws.ev[k]) ZP_CUDA(cudaEventCreate(&ws.ev[k]));
        ZP_CUDA(cudaEventRecord(ws.ev[k], st));
    };
    ZP_CUDA(cudaMemsetAsync(ws.cursor.p, 0, wb * sizeof(uint32_t), st));
    mark(0);
    if (n) {
        ZP_LAUNCH(msm_digits_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, scalars, n, cfg.c, cfg.nwin, cfg.nbuckets,
                  ws.digits.p, ws.cursor.p);
    }
    mark(1);
    ZP_LAUNCH(msm_scan_kernel, dim3(1), dim3(1024), 0, st, ws.cursor.p, ws.start.p, wb);
    mark(2);
    if (n) {
        ZP_LAUNCH(msm_scatter_kernel, dim3((unsigned)((n + 255) / 256), cfg.nwin), dim3(256), 0, st, ws.digits.p, n, cfg.nwin,
                  cfg.nbuckets, ws.cursor.p, ws.sorted.p);
    }
    ZP_LAUNCH(msm_segcount_kernel, dim3((unsigned)((wb + 255) / 256)), dim3(256), 0, st, ws.start.p, ws.cursor.p, wb,
              (uint32_t)ws.seg, ws.seg_cnt.p);
    ZP_LAUNCH(msm_scan_kernel, dim3(1), dim3(1024), 0, st, ws.seg_cnt.p, ws.seg_start.p, wb);
    mark(3);
    ZP_LAUNCH(msm_accumulate_kernel, dim3((unsigned)((ws.max_segs + 127) / 128)), dim3(128), 0, st, points, ws.sorted.p,
              ws.start.p, ws.cursor.p, ws.seg_start.p, wb, (uint32_t)ws.seg, ws.segs.p);
    mark(4);
    int groups = MSM_REDUCE_GROUPS;
    while (cfg.nbuckets / groups < 1) groups >>= 1;
    int bg = cfg.nbuckets / groups;
    int T = bg < 128 ? bg : 128;
    ZP_LAUNCH(msm_reduce_kernel, dim3(cfg.nwin * groups), dim3(T), (size_t)T * sizeof(xyzz_t), st, ws.segs.p, ws.seg_start.p,
              cfg.nbuckets, groups, ws.partial.p);
    mark(5);
    ZP_CUDA(cudaMemcpyAsync(ws.partial_host.data(), ws.partial.p, (size_t)cfg.nwin * groups * sizeof(xyzz_t),
                            cudaMemcpyDeviceToHost, st));
}

host::G1 msm_collect(MsmWorkspace& ws, const MsmConfig& cfg, cudaStream_t st) {
    ZP_CUDA(cudaStreamSynchronize(st));
    if (ws.timing && ws.ev[5]) {
        for (int k = 0; k < 5; k++) {
            float ms = 0;
            ZP_CUDA(cudaEventElapsedTime(&ms, ws.ev[k], ws.ev[k + 1]));
            ws.last_ms[k] = ms;
        }
    }
    int groups = MSM_REDUCE_GROUPS;
    while (cfg.nbuckets / groups < 1) groups >>= 1;
    host::G1 total = host::G1::infinity();
    for (int w = cfg.nwin - 1; w >= 0; w--) {
        for (int b = 0; b < cfg.c; b++) total.dbl_inplace();
        for (int g = 0; g < groups; g++) total.add(host::G1::from_dev(ws.partial_host[(size_t)w * groups + g]));
    }
    return total;
}

}  // namespace zp
