// Pippenger MSM over BLS12-381 G1 for sm_100a — see msm.cuh / DESIGN.md §MSM.
//
// Pipeline of one batch of k <= 8 scalar vectors over the same points (all on one stream; host round trips: the top of the
// inversion tree of each batch-affine round and the k final sums):
//   1. digits    : Montgomery scalar -> canonical -> signed c-bit digits, per-(member, window / set, bucket) histogram
//   2. scan      : exclusive scan of the histogram (bucket start offsets), multi-CTA
//   3. scatter   : counting-sort the point indices into bucket runs (atomic cursor per bucket)
//   4. rounds    : batch-affine pairwise additions inside every run (msm_affine.cuh), 4 rounds by default; every round
//                  writes its partial sums as x[] / y[] arrays that the next round reads sequentially
//   5. accumulate: what is left (1/16 of the entries): one thread per work segment, XYZZ mixed additions, persistent
//                  threads; long buckets are split so no digit distribution serialises (and folded back from a work list)
//   6. reduce    : row / column sums of the bucket matrix, then weight * sum and a tree (msm_rowcol_kernel, msm_weighted_kernel)
//   host         : with precomputed window tables one XYZZ point per member; otherwise Horner over the window sums; to affine
// Multi-GPU (precomputed tables only): ranks split the BUCKETS, not the points.  A launch with cfg.bucket_lg = log2 G keeps
// only the digits whose bucket g satisfies g mod G == bucket_rank and numbers them locally j = g / G; every rank walks all
// scalars (cheap, HBM-bound) but accumulates 1/G of the bucket entries at the SAME window size and bucket load as one GPU
// (point-range slices force c down — 16 windows instead of 13 at 2^19 points — and starve the batch-affine rounds).  The
// split is CYCLIC because the short top window only produces small digits: contiguous ranges would hand all of its
// entries to rank 0 (+67 % load, measured as a straggler at G = 8).  The reduction returns W = sum (j + 1) B_j and the plain
// sum T of the rank's buckets; the true weight of local bucket j is j G + bucket_rank + 1, so the rank's share is
// G W + (bucket_rank + 1 - G) T  (two short scalar multiplications on the host).
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
    cfg.nsets = cfg.nwin;
    cfg.tab_stride = 0;
    cfg.pt_stride = (uint32_t)sizeof(affine_t);
    return cfg;
}

MsmConfig msm_config_precomp(size_t n, size_t tab_stride) {
    MsmConfig cfg;
    int lg = ilog2(n < 2 ? 2 : n);
    int c = lg - 2;
    if (c < 12) c = 12;
    if (c > 20) c = 20;
    cfg.c = c;
    cfg.nwin = (256 + c - 1) / c;
    cfg.nbuckets = 1 << (c - 1);
    cfg.nsets = 1;
    cfg.tab_stride = tab_stride;
    cfg.pt_stride = (uint32_t)sizeof(affine_pad_t);
    return cfg;
}

static const int SCAN_TILE_FWD = 2048;
// bucket matrix of the reduction (see msm_rowcol_kernel): 2^lw2 columns
static int msm_reduce_lw2(const MsmConfig& cfg) { return (ilog2((size_t)cfg.nbuckets) + 1) / 2; }  // W2 >= W1; c / 2 for 2^(c-1) buckets
static int msm_reduce_entries(const MsmConfig& cfg) { return (cfg.nbuckets >> msm_reduce_lw2(cfg)) + (1 << msm_reduce_lw2(cfg)); }
static int msm_reduce_groups(const MsmConfig& cfg) { return (msm_reduce_entries(cfg) + 127) / 128; }

void MsmWorkspace::reserve(size_t n, const MsmConfig& cfg, int nbatch) {
    size_t wn = (size_t)nbatch * cfg.nwin * n, wb = (size_t)nbatch * cfg.nsets * cfg.nbuckets;
    if (digits.n < wn) digits.alloc(wn);
    if (sorted.n < wn) sorted.alloc(wn);
    if (start.n < wb + 1) start.alloc(wb + 1);
    if (cursor.n < wb) cursor.alloc(wb);
    if (seg_start.n < wb + 1) seg_start.alloc(wb + 1);
    if (seg_cnt.n < wb) seg_cnt.alloc(wb);
    if (tile_sum.n < wb / SCAN_TILE_FWD + 2) tile_sum.alloc(wb / SCAN_TILE_FWD + 2);
    if (!sm_count) {
        int dev = 0;
        cudaDeviceProp prop;
        ZP_CUDA(cudaGetDevice(&dev));
        ZP_CUDA(cudaGetDeviceProperties(&prop, dev));
        sm_count = prop.multiProcessorCount;
        const char* v = getenv("ZP_ACC_VARIANT");
        if (v) acc_variant = atoi(v);
    }
    if (ba_rounds == 0) {
        const char* br = getenv("ZP_MSM_BA_ROUNDS");
        // default 4 rounds: measured 25.4 -> 19.9 ms at 2^22, 17.5 ms per MSM in a batch of 4 (profiles/r01b_msm_*.log);
        // 0 disables
        ba_rounds = br ? atoi(br) : 4;
        ba_rounds_forced = br != nullptr;
        if (ba_rounds <= 0) ba_rounds = -1;  // disabled
        const char* bm = getenv("ZP_MSM_BA_MIN_LOG");
        if (bm) ba_min_entries = (size_t)1 << atoi(bm);
    }
    if (!counter.p) counter.alloc(2);  // [0] segment counter of the accumulate kernel, [1] length of fold_list
    if (fold_list.n < wb) fold_list.alloc(wb);
    size_t nsets = (size_t)nbatch * cfg.nsets;
    size_t np = nsets * msm_reduce_groups(cfg);
    if (partial.n < np) partial.alloc(np);
    if (final_sums.n < 2 * nsets) final_sums.alloc(2 * nsets);
    if (rowcol.n < nsets * msm_reduce_entries(cfg)) rowcol.alloc(nsets * msm_reduce_entries(cfg));
    if (np < 2 * nsets) np = 2 * nsets;
    if (partial_host.size() < np) partial_host.resize(np);
}

// blockIdx.y = member of the batch (several scalar vectors over the same points, one bucket-set group each)
__global__ void __launch_bounds__(256) msm_digits_kernel(MsmBatch batch, size_t n, int c, int nwin, int nbuckets, int bucket_lg,
                                                         uint32_t bucket_rank, int one_set, uint32_t* __restrict__ digits, uint32_t* __restrict__ hist) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fr_t* __restrict__ scalars = batch.s[blockIdx.y];
    if (digits) digits += (size_t)blockIdx.y * nwin * n;  // null: histogram only, the scatter recomputes the digits
    hist += (size_t)blockIdx.y * (one_set ? 1 : nwin) * nbuckets;
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
        // bucket share of this launch: global bucket g = d - 1 is ours iff g mod 2^bucket_lg == bucket_rank; local index g >> bucket_lg
        if (d && bucket_lg) {
            const uint32_t g = d - 1;
            d = (g & ((1u << bucket_lg) - 1)) == bucket_rank ? (g >> bucket_lg) + 1 : 0;
        }
        if (digits) digits[(size_t)w * n + i] = d | (neg << 31);
        if (d) atomicAdd(&hist[(one_set ? 0 : (size_t)w * nbuckets) + d - 1], 1u);
    }
}

// Exclusive scan of cnt[0..m) in three small launches (tile sums, scan of tile sums, tile-local scan):
// start[0..m] = exclusive scan; cnt[i] <- start[i] (becomes the scatter cursor).
static const int SCAN_TILE = 2048;  // elements per CTA (256 threads x 8)
__global__ void __launch_bounds__(256) msm_scan_tiles_kernel(const uint32_t* __restrict__ cnt, size_t m, uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t red[256];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * 8;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (base + k < m) s += cnt[base + k];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
        if ((int)threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = red[0];
}
// single CTA: exclusive scan of up to 1024 * chunk tile sums in place; total written to *total_out
__global__ void __launch_bounds__(1024) msm_scan_sums_kernel(uint32_t* __restrict__ tile_sum, size_t ntiles, uint32_t* __restrict__ total_out) {
    __shared__ uint32_t part[1024];
    const int t = threadIdx.x;
    size_t chunk = (ntiles + 1023) / 1024;
    size_t lo = (size_t)t * chunk, hi = lo + chunk < ntiles ? lo + chunk : ntiles;
    uint32_t s = 0;
    for (size_t i = lo; i < hi; i++) s += tile_sum[i];
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        uint32_t v = (t >= d) ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - s;
    for (size_t i = lo; i < hi; i++) {
        uint32_t c = tile_sum[i];
        tile_sum[i] = run;
        run += c;
    }
    if (t == 1023) *total_out = part[1023];
}
__global__ void __launch_bounds__(256) msm_scan_apply_kernel(uint32_t* __restrict__ cnt, uint32_t* __restrict__ start, size_t m,
                                                             const uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t part[256];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * 8;
    uint32_t v[8], s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        v[k] = base + k < m ? cnt[base + k] : 0;
        s += v[k];
    }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        uint32_t x = ((int)threadIdx.x >= d) ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += x;
        __syncthreads();
    }
    uint32_t run = tile_sum[blockIdx.x] + part[threadIdx.x] - s;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (base + k < m) {
            start[base + k] = run;
            cnt[base + k] = run;
        }
        run += v[k];
    }
}
void u32_exclusive_scan(uint32_t* cnt, uint32_t* start, size_t m, uint32_t* tile_sum, cudaStream_t st);
static void msm_scan(uint32_t* cnt, uint32_t* start, size_t m, uint32_t* tile_sum, cudaStream_t st) {
    u32_exclusive_scan(cnt, start, m, tile_sum, st);
}
void u32_exclusive_scan(uint32_t* cnt, uint32_t* start, size_t m, uint32_t* tile_sum, cudaStream_t st) {
    size_t ntiles = (m + SCAN_TILE - 1) / SCAN_TILE;
    ZP_LAUNCH(msm_scan_tiles_kernel, dim3((unsigned)ntiles), dim3(256), 0, st, cnt, m, tile_sum);
    ZP_LAUNCH(msm_scan_sums_kernel, dim3(1), dim3(1024), 0, st, tile_sum, ntiles, start + m);
    ZP_LAUNCH(msm_scan_apply_kernel, dim3((unsigned)ntiles), dim3(256), 0, st, cnt, start, m, tile_sum);
}

__global__ void __launch_bounds__(256) msm_scatter_kernel(const uint32_t* __restrict__ digits, size_t n, int nwin, int nbuckets,
                                                          size_t tab_stride, uint32_t* __restrict__ cursor,
                                                          uint32_t* __restrict__ sorted) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = blockIdx.y % nwin, b = blockIdx.y / nwin;  // window, batch member
    if (i >= n) return;
    uint32_t dg = digits[(size_t)blockIdx.y * n + i];
    uint32_t d = dg & 0x7fffffffu;
    if (!d) return;
    const size_t set = tab_stride ? (size_t)b : (size_t)b * nwin + w;
    uint32_t pos = atomicAdd(&cursor[set * nbuckets + d - 1], 1u);
    // with a precomputed table the entry addresses row w of the table: 2^(c w) * P_i
    sorted[pos] = (uint32_t)(tab_stride ? (size_t)w * tab_stride + i : i) | (dg & 0x80000000u);
}

// Scatter that recomputes the signed digits from the scalars instead of reading the [nwin][n] digit array (which is then
// never written): 32 B read per scalar instead of 4 B written + 4 B read per (scalar, window) — with bucket shares 7/8 of
// those digits are dropped anyway.  One thread per scalar, the same recoding as msm_digits_kernel.
__global__ void __launch_bounds__(256) msm_scatter_fused_kernel(MsmBatch batch, size_t n, int c, int nwin, int nbuckets, int bucket_lg,
                                                                uint32_t bucket_rank, size_t tab_stride,
                                                                uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = blockIdx.y;
    fr_t s = load_fr(&batch.s[b][i]).from_mont();
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
        if (d && bucket_lg) {
            const uint32_t g = d - 1;
            d = (g & ((1u << bucket_lg) - 1)) == bucket_rank ? (g >> bucket_lg) + 1 : 0;
        }
        if (!d) continue;
        const size_t set = tab_stride ? (size_t)b : (size_t)b * nwin + w;
        uint32_t pos = atomicAdd(&cursor[set * nbuckets + d - 1], 1u);
        sorted[pos] = (uint32_t)(tab_stride ? (size_t)w * tab_stride + i : i) | (neg << 31);
    }
}

// point number idx of an array whose elements are `stride` bytes apart (96: affine_t, 128: affine_pad_t)
ZP_D const affine_t* point_at(const affine_t* base, uint32_t idx, uint32_t stride) {
    return reinterpret_cast<const affine_t*>(reinterpret_cast<const unsigned char*>(base) + (size_t)idx * stride);
}
// Where a kernel reads its points from: an array of structures `stride` bytes apart (the SRS, a window table, a caller's
// points) or, for the partial sums the batch-affine rounds materialise, two arrays — x[i] at base, y[i] `ysep` elements later.
struct PointSrc {
    const affine_t* base;
    uint32_t stride;
    size_t ysep;  // 0: array of structures
};
ZP_D fq_t point_x(const PointSrc& s, uint32_t i) {
    return s.ysep ? load_fq(&reinterpret_cast<const fq_t*>(s.base)[i]) : load_fq(&point_at(s.base, i, s.stride)->x);
}
ZP_D fq_t point_y(const PointSrc& s, uint32_t i) {
    return s.ysep ? load_fq(&reinterpret_cast<const fq_t*>(s.base)[s.ysep + i]) : load_fq(&point_at(s.base, i, s.stride)->y);
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

// Segment descriptors (first / one-past-last index into `sorted`), one thread per bucket.
__global__ void __launch_bounds__(256) msm_segdesc_kernel(const uint32_t* __restrict__ start, const uint32_t* __restrict__ endp,
                                                          const uint32_t* __restrict__ seg_start, size_t nb, uint32_t seg,
                                                          uint2* __restrict__ desc) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint32_t s0 = seg_start[b], s1 = seg_start[b + 1], k = start[b], e = endp[b];
    for (uint32_t s = s0; s < s1; s++) {
        uint2 d;
        d.x = k;
        d.y = k + seg < e ? k + seg : e;
        desc[s] = d;
        k += seg;
    }
}

// XYZZ mixed additions over the work segments.  Persistent threads: every lane pulls its next segment
// from a global counter as soon as it finishes one, so all 32 lanes of a warp keep executing the same
// point-addition body (no tail divergence from unequal bucket sizes).
template <int MINBLOCKS>
__global__ void __launch_bounds__(128, MINBLOCKS) msm_accumulate_kernel(const PointSrc ps,
                                                                        const uint32_t* __restrict__ sorted,
                                                                        const uint2* __restrict__ desc,
                                                                        const uint32_t* __restrict__ nseg_ptr,
                                                                        uint32_t* __restrict__ counter, xyzz_t* __restrict__ segs) {
    const uint32_t nseg = *nseg_ptr;
    uint32_t s = atomicAdd(counter, 1u);
    if (s >= nseg) return;
    uint2 d = desc[s];
    uint32_t k = d.x, k1 = d.y;
    xyzz_t acc = xyzz_t::infinity();
    while (true) {
        if (k >= k1) {
            store_xyzz(&segs[s], acc);
            s = atomicAdd(counter, 1u);
            if (s >= nseg) break;
            d = desc[s];
            k = d.x;
            k1 = d.y;
            acc = xyzz_t::infinity();
        }
        // (software-pipelining the gather one iteration ahead was measured slower: 32.3 vs 30.1 ms — the loop is
        // multiplier-bound, the extra 24 live registers cost more than the hidden latency)
        uint32_t e = sorted ? sorted[k] : k;  // after batch-affine rounds the run IS the point array
        affine_t p;
        p.x = point_x(ps, e & 0x7fffffffu);
        p.y = point_y(ps, e & 0x7fffffffu);
        if (e >> 31) p.y = p.y.neg();
        acc.add_affine(p.x, p.y);
        k++;
    }
}

// Buckets split into a few work segments (the common case on small / sharded inputs) are folded by one THREAD each
// (all lanes busy); long runs (skewed digit distributions) are appended to a work list and folded by one warp each in
// msm_fold_kernel, launched with a fixed grid (one warp per BUCKET used to cost 0.7 ms per pipeline in CTAs that returned
// at once).
static const uint32_t FOLD_SERIAL_MAX = 8;
__global__ void __launch_bounds__(128) msm_fold_small_kernel(xyzz_t* __restrict__ segs, const uint32_t* __restrict__ seg_start, size_t nb,
                                                             uint32_t* __restrict__ long_list, uint32_t* __restrict__ long_count) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint32_t s0 = seg_start[b], s1 = seg_start[b + 1];
    if (s1 - s0 < 2) return;
    if (s1 - s0 > FOLD_SERIAL_MAX) {
        long_list[atomicAdd(long_count, 1u)] = (uint32_t)b;
        return;
    }
    xyzz_t acc = load_xyzz(&segs[s0]);
    for (uint32_t s = s0 + 1; s < s1; s++) acc.add(load_xyzz(&segs[s]));
    store_xyzz(&segs[s0], acc);
}

// The listed buckets are folded back by one warp each (lanes stride over the segment sums, then a shared-memory tree);
// afterwards segs[seg_start[b]] holds the whole bucket.
__global__ void __launch_bounds__(128) msm_fold_kernel(xyzz_t* __restrict__ segs, const uint32_t* __restrict__ seg_start,
                                                       const uint32_t* __restrict__ long_list, const uint32_t* __restrict__ long_count) {
    __shared__ xyzz_t sm[128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t count = *long_count;
    xyzz_t* w = sm + warp * 32;
    for (uint32_t i = blockIdx.x * 4 + warp; i < count; i += gridDim.x * 4) {  // warp-uniform
        const uint32_t b = long_list[i];
        const uint32_t s0 = seg_start[b], s1 = seg_start[b + 1];
        xyzz_t acc = xyzz_t::infinity();
        for (uint32_t s = s0 + lane; s < s1; s += 32) acc.add(load_xyzz(&segs[s]));
        w[lane] = acc;
        __syncwarp();
        for (int d = 16; d >= 1; d >>= 1) {
            if (lane < d) {
                xyzz_t a = w[lane];
                a.add(w[lane + d]);
                w[lane] = a;
            }
            __syncwarp();
        }
        if (lane == 0) store_xyzz(&segs[s0], w[0]);
        __syncwarp();
    }
}

ZP_D xyzz_t load_bucket(const xyzz_t* __restrict__ segs, const uint32_t* __restrict__ seg_start, size_t b) {
    uint32_t s0 = seg_start[b], s1 = seg_start[b + 1];
    if (s0 == s1) return xyzz_t::infinity();
    return load_xyzz(&segs[s0]);
}

// Bucket reduction  sum_j (j + 1) * B_j  of one bucket set in two short steps instead of one long running sum.
// With j = j1 * W2 + j2 (W2 = 2^lw2 columns, W1 rows):
//     sum_j (j + 1) B_j  =  sum_{j1} (j1 * W2) * R_{j1}  +  sum_{j2} (j2 + 1) * C_{j2}
// where R_{j1} / C_{j2} are the plain row / column sums of the bucket matrix.  Step 1 (msm_rowcol_kernel) computes the
// W1 + W2 plain sums, one CTA each (throughput-bound, 2 additions per bucket in total); step 2 (msm_weighted_kernel)
// multiplies each of them by its <= c-bit weight and tree-adds.  The dependent chain is ~30 additions long instead of
// the ~100 of a 128-thread running sum over 2^19 buckets, which was latency-bound at 1.9 ms per MSM.
// rc[set][0 .. W1) = R_{j1} (entry 0 has weight 0 and is never read), rc[set][W1 .. W1 + W2) = C_{j2}.
__global__ void __launch_bounds__(128) msm_rowcol_kernel(const xyzz_t* __restrict__ segs, const uint32_t* __restrict__ seg_start,
                                                         int nbuckets, int lw2, xyzz_t* __restrict__ rc) {
    __shared__ xyzz_t sm[128];
    const int W2 = 1 << lw2, W1 = nbuckets >> lw2, E = W1 + W2;
    const int set = blockIdx.x / E, o = blockIdx.x % E;
    const size_t B0 = (size_t)set * nbuckets;
    xyzz_t acc = xyzz_t::infinity();
    if (o < W1) {
        if (o == 0) return;  // weight 0
        for (int m = threadIdx.x; m < W2; m += blockDim.x) acc.add(load_bucket(segs, seg_start, B0 + (size_t)o * W2 + m));
    } else {
        const int j2 = o - W1;
        for (int m = threadIdx.x; m < W1; m += blockDim.x) acc.add(load_bucket(segs, seg_start, B0 + (size_t)m * W2 + j2));
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int d = blockDim.x >> 1; d >= 1; d >>= 1) {
        if ((int)threadIdx.x < d) {
            xyzz_t a = sm[threadIdx.x];
            a.add(sm[threadIdx.x + d]);
            sm[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) store_xyzz(&rc[blockIdx.x], sm[0]);
}

// partial[set * G + g] = sum over the g-th 128 entries of rc[set] of weight * entry  (G = ceil((W1 + W2) / 128))
__global__ void __launch_bounds__(128) msm_weighted_kernel(const xyzz_t* __restrict__ rc, int nbuckets, int lw2, int groups,
                                                           xyzz_t* __restrict__ partial) {
    __shared__ xyzz_t sm[128];
    const int W2 = 1 << lw2, W1 = nbuckets >> lw2, E = W1 + W2;
    const int set = blockIdx.x / groups, g = blockIdx.x % groups;
    const int o = g * 128 + threadIdx.x;
    xyzz_t sum = xyzz_t::infinity();
    if (o < E && o != 0) {
        const uint32_t wgt = o < W1 ? (uint32_t)o << lw2 : (uint32_t)(o - W1 + 1);
        const xyzz_t v = load_xyzz(&rc[(size_t)set * E + o]);
        for (int bit = 31 - __clz(wgt); bit >= 0; bit--) {
            sum.dbl_inplace();
            if ((wgt >> bit) & 1) sum.add(v);
        }
    }
    sm[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 64; d >= 1; d >>= 1) {
        if ((int)threadIdx.x < d) {
            xyzz_t a = sm[threadIdx.x];
            a.add(sm[threadIdx.x + d]);
            sm[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) store_xyzz(&partial[blockIdx.x], sm[0]);
}

}  // namespace zp
#include "msm_affine.cuh"
namespace zp {

// ---- precomputed window table T[w][i] = 2^(c w) * P_i, one row from the previous one.
// Row w is c doublings of row w - 1 in XYZZ (msm_table_dbl_kernel), then back to affine with ONE shared inversion per row:
// z_i = ZZ_i * ZZZ_i goes through the batch inversion of the batch-affine rounds (3 products per point + one host round
// trip) instead of a 570-product Fermat inversion per point — 735 -> 170 products per table entry, the 6.5 GiB table of a
// 2^22-point SRS builds in a quarter of the time (it is most of a context's cold first call).
__global__ void __launch_bounds__(128) msm_table_copy_kernel(affine_pad_t* __restrict__ dst, const affine_t* __restrict__ src, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_fq(&dst[i].x, load_fq(&src[i].x));
    store_fq(&dst[i].y, load_fq(&src[i].y));
}
__global__ void __launch_bounds__(128) msm_table_dbl_kernel(const affine_pad_t* __restrict__ prev, size_t n, int c,
                                                            xyzz_t* __restrict__ tmp, fq_t* __restrict__ z) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fq_t x = load_fq(&prev[i].x), y = load_fq(&prev[i].y);
    xyzz_t a;
    a.set_double_affine(x, y);
    for (int k = 1; k < c; k++) a.dbl_inplace();
    store_xyzz(&tmp[i], a);
    // the group has prime order, so 2^c * P is finite for a finite P on the curve; anything else (a caller's (0, 0)) ends
    // with ZZ = 0 and must not poison the shared inversion: it takes part as 1 and is written back as (0, 0)
    fq_t zz = a.ZZ * a.ZZZ;
    store_fq(&z[i], zz.is_zero() ? fq_t::one() : zz);
}
__global__ void __launch_bounds__(128) msm_table_affine_kernel(affine_pad_t* __restrict__ dst, const xyzz_t* __restrict__ tmp,
                                                               const fq_t* __restrict__ zinv, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const xyzz_t a = load_xyzz(&tmp[i]);
    fq_t x = fq_t::zero(), y = fq_t::zero();
    if (!a.is_inf()) {
        const fq_t inv = load_fq(&zinv[i]);  // 1 / (ZZ * ZZZ)
        x = a.X * (inv * a.ZZZ);             // X / ZZ
        y = a.Y * (inv * a.ZZ);              // Y / ZZZ
    }
    store_fq(&dst[i].x, x);
    store_fq(&dst[i].y, y);
}
void msm_build_table(affine_pad_t* dst, const affine_t* src, size_t n, int c, int nwin, cudaStream_t st) {
    if (!n) return;
    const dim3 grid((unsigned)((n + 127) / 128)), block(128);
    ZP_LAUNCH(msm_table_copy_kernel, grid, block, 0, st, dst, src, n);
    if (nwin <= 1) return;
    DevBuf<xyzz_t> tmp(n);
    DevBuf<fq_t> z(n + n / (BI_CH - 1) + 64);  // the values, followed by the upper levels of the inversion tree
    PinnedFq root;
    for (int w = 1; w < nwin; w++) {
        ZP_LAUNCH(msm_table_dbl_kernel, grid, block, 0, st, dst + (size_t)(w - 1) * n, n, c, tmp.p, z.p);
        fq_batch_inverse(z.p, n, z.p + n, root.get(), st);
        ZP_LAUNCH(msm_table_affine_kernel, grid, block, 0, st, dst + (size_t)w * n, tmp.p, z.p, n);
    }
    ZP_CUDA(cudaStreamSynchronize(st));  // tmp / z / root die with this frame
}

// final[set] = sum_g partial[set * groups + g]   (one CTA per bucket set)
__global__ void __launch_bounds__(128) msm_final_kernel(const xyzz_t* __restrict__ partial, int groups, xyzz_t* __restrict__ final_out) {
    __shared__ xyzz_t sm[128];
    const xyzz_t* P = partial + (size_t)blockIdx.x * groups;
    xyzz_t acc = xyzz_t::infinity();
    for (int g = threadIdx.x; g < groups; g += blockDim.x) acc.add(load_xyzz(&P[g]));
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int d = blockDim.x >> 1; d >= 1; d >>= 1) {
        if ((int)threadIdx.x < d) {
            xyzz_t a = sm[threadIdx.x];
            a.add(sm[threadIdx.x + d]);
            sm[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) store_xyzz(&final_out[blockIdx.x], sm[0]);
}

// total[set] = sum of the W2 column sums of the set (all buckets once)
__global__ void __launch_bounds__(128) msm_plain_total_kernel(const xyzz_t* __restrict__ rc, int W1, int W2, xyzz_t* __restrict__ total) {
    __shared__ xyzz_t sm[128];
    const xyzz_t* C = rc + (size_t)blockIdx.x * (W1 + W2) + W1;
    xyzz_t acc = xyzz_t::infinity();
    for (int j = threadIdx.x; j < W2; j += blockDim.x) acc.add(load_xyzz(&C[j]));
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int d = blockDim.x >> 1; d >= 1; d >>= 1) {
        if ((int)threadIdx.x < d) {
            xyzz_t a = sm[threadIdx.x];
            a.add(sm[threadIdx.x + d]);
            sm[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) store_xyzz(&total[blockIdx.x], sm[0]);
}

static void msm_launch_impl(MsmWorkspace& ws, const MsmConfig& cfg, const affine_t* points, const MsmBatch& batch, int nbatch, size_t n,
                            bool allow_ba, cudaStream_t st) {
    if (nbatch < 1 || nbatch > MSM_MAX_BATCH) throw std::runtime_error("msm: batch size out of range");
    ws.reserve(n, cfg, nbatch);
    const int nsets = nbatch * cfg.nsets;
    const size_t wb = (size_t)nsets * cfg.nbuckets, wn = (size_t)nbatch * cfg.nwin * n;
    if (wn >= ((size_t)1 << 31)) throw std::runtime_error("msm: batch too large for 31-bit entry indices");
    if (cfg.tab_stride && (size_t)cfg.nwin * cfg.tab_stride >= ((size_t)1 << 31))
        throw std::runtime_error("msm: precomputed table too large for 31-bit indices");
    auto mark = [&](int k) {
        if (!ws.timing) return;
        if (!ws.ev[k]) ZP_CUDA(cudaEventCreate(&ws.ev[k]));
        ZP_CUDA(cudaEventRecord(ws.ev[k], st));
    };
    ws.last_points = points;
    ws.last_batch = batch;
    ws.last_nbatch = nbatch;
    ws.last_n = n;
    ZP_CUDA(cudaMemsetAsync(ws.cursor.p, 0, wb * sizeof(uint32_t), st));
    mark(0);
    // no digit array, the scatter recomputes the digits (msm_scatter_fused_kernel): a win only with bucket shares, where most
    // digits are dropped — one rank's share of 8 at 2^22 x 4: digits + scatter 1.55 -> 1.18 ms; whole MSM: 4.25 -> 4.66 ms
    // (profiles/r02s_msm_fused_scatter.log).  ZP_MSM_FUSED_SCATTER=0|1 overrides.
    static const int fused_env = getenv("ZP_MSM_FUSED_SCATTER") ? atoi(getenv("ZP_MSM_FUSED_SCATTER")) : -1;
    const bool fused_scatter = fused_env >= 0 ? fused_env != 0 : cfg.bucket_lg > 0;
    if (n) {
        ZP_LAUNCH(msm_digits_kernel, dim3((unsigned)((n + 255) / 256), nbatch), dim3(256), 0, st, batch, n, cfg.c, cfg.nwin,
                  cfg.nbuckets, cfg.bucket_lg, cfg.bucket_rank, cfg.tab_stride ? 1 : 0, fused_scatter ? nullptr : ws.digits.p, ws.cursor.p);
    }
    mark(1);
    msm_scan(ws.cursor.p, ws.start.p, wb, ws.tile_sum.p, st);
    mark(2);
    if (n && fused_scatter) {
        ZP_LAUNCH(msm_scatter_fused_kernel, dim3((unsigned)((n + 255) / 256), nbatch), dim3(256), 0, st, batch, n, cfg.c, cfg.nwin,
                  cfg.nbuckets, cfg.bucket_lg, cfg.bucket_rank, cfg.tab_stride, ws.cursor.p, ws.sorted.p);
    } else if (n) {
        ZP_LAUNCH(msm_scatter_kernel, dim3((unsigned)((n + 255) / 256), cfg.nwin * nbatch), dim3(256), 0, st, ws.digits.p, n, cfg.nwin,
                  cfg.nbuckets, cfg.tab_stride, ws.cursor.p, ws.sorted.p);
    }
    mark(3);
    // ---- batch-affine pre-reduction rounds
    const uint32_t *run_begin = ws.start.p, *run_end = ws.cursor.p, *entries = ws.sorted.p;
    PointSrc ps{points, cfg.pt_stride, 0};  // the batch-affine rounds materialise their partial sums as x[] / y[] arrays
    size_t est = wn;  // upper bound on the bucket entries still to be added
    const bool bucket_slice = cfg.bucket_lg > 0;
    if (bucket_slice) {
        // only the digits of this rank's bucket slice were kept: size the rounds by what is really there (one 4-byte read)
        uint32_t total = 0;
        ZP_CUDA(cudaMemcpyAsync(&total, ws.start.p + wb, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        ZP_CUDA(cudaStreamSynchronize(st));
        est = total;
    }
    const size_t wn_eff = est;
    int rounds = (allow_ba && wn_eff >= ws.ba_min_entries) ? ws.ba_rounds : 0;
    // leave >= 4 entries per bucket on average for the XYZZ pass (each round has a fixed cost of ~0.5 ms)
    if (!ws.ba_rounds_forced)
        while (rounds > 0 && ((wn_eff / wb) >> rounds) < 4) rounds--;
    ws.ba_used = rounds > 0;
    if (rounds > 0) {
        size_t cap0 = wn_eff / 2 + wb;
        size_t up0 = (cap0 / (BA_K * BA_T) + 1) * BA_T;  // leaf groups (level-1 nodes of the inversion tree)
        if (ws.ba_pts[0].n < cap0) ws.ba_pts[0].alloc(cap0);
        if (rounds > 1 && ws.ba_pts[1].n < cap0 / 2 + wb) ws.ba_pts[1].alloc(cap0 / 2 + wb);
        if (ws.ba_pre.n < cap0) ws.ba_pre.alloc(cap0);
        if (ws.ba_den.n < up0 + up0 / (BI_CH - 1) + 64) ws.ba_den.alloc(up0 + up0 / (BI_CH - 1) + 64);
        if (ws.ba_src.n < cap0) ws.ba_src.alloc(cap0);
        if (ws.ba_cnt.n < wb) ws.ba_cnt.alloc(wb);
        for (int k = 0; k < 2; k++)
            if (ws.ba_rs[k].n < wb + 1) ws.ba_rs[k].alloc(wb + 1);
        if (!ws.ba_flag.p) ws.ba_flag.alloc(2);
        ZP_CUDA(cudaMemsetAsync(ws.ba_flag.p, 0, 2 * sizeof(uint32_t), st));
        const bool stats = ws.timing && rounds <= MsmWorkspace::BA_MAX_ROUNDS;
        for (int r = 0; r < rounds; r++) {
            size_t cap = est / 2 + wb;
            uint32_t* rs = ws.ba_rs[r & 1].p;
            // this round's partial sums: x[0 .. cap) then y[0 .. cap) in the ping-pong buffer (cap * 96 bytes)
            fq_t* out_x = reinterpret_cast<fq_t*>(ws.ba_pts[r & 1].p);
            fq_t* out_y = out_x + cap;
            unsigned nblk = (unsigned)((cap + BA_K * BA_T - 1) / (BA_K * BA_T));
            size_t m = (size_t)nblk * BA_T;
            ZP_LAUNCH(ba_pair_count_kernel, dim3((unsigned)((wb + 255) / 256)), dim3(256), 0, st, run_begin, run_end, wb, ws.ba_cnt.p);
            msm_scan(ws.ba_cnt.p, rs, wb, ws.tile_sum.p, st);
            ZP_LAUNCH(ba_slots_kernel, dim3((unsigned)((wb * 8 + 255) / 256)), dim3(256), 0, st, run_begin, run_end, rs, wb,
                      ws.ba_src.p);
            ZP_LAUNCH(ba_up0_kernel, dim3(nblk), dim3(BA_T), 0, st, ws.ba_src.p, rs + wb, cap, entries, ps, ws.ba_pre.p, ws.ba_den.p,
                      ws.ba_flag.p);
            fq_batch_inverse(ws.ba_den.p, m, ws.ba_den.p + m, ws.ba_root.get(), st);
            if (stats) {
                if (!ws.down_ev[2 * r]) {
                    ZP_CUDA(cudaEventCreate(&ws.down_ev[2 * r]));
                    ZP_CUDA(cudaEventCreate(&ws.down_ev[2 * r + 1]));
                }
                ZP_CUDA(cudaEventRecord(ws.down_ev[2 * r], st));
            }
            ZP_LAUNCH(ba_down0_kernel, dim3(nblk), dim3(BA_T), 0, st, ws.ba_src.p, rs + wb, cap, entries, ps, ws.ba_pre.p,
                      ws.ba_den.p, out_x, out_y);
            if (stats) ZP_CUDA(cudaEventRecord(ws.down_ev[2 * r + 1], st));
            run_begin = rs;
            run_end = rs + 1;
            entries = nullptr;
            ps = PointSrc{reinterpret_cast<const affine_t*>(out_x), 0u, cap};
            est = cap;
        }
    }
    // ---- work segments over the (remaining) runs.  Length: at most 2x the mean bucket load (typical buckets are one
    // segment) but short enough that every persistent thread gets >= 8 segments; never below 32 (16 after pre-reduction).
    {
        size_t mean = (est + wb - 1) / wb;
        size_t nthreads = (size_t)ws.sm_count * 3 * 128;
        size_t balanced = est / (nthreads * 8);
        size_t seg = 2 * mean;
        if (balanced < seg) seg = balanced;
        size_t floor_len = rounds > 0 ? 16 : 32;
        if (seg < floor_len) seg = floor_len;
        ws.seg = seg;
        size_t need = est / seg + wb + 1;
        if (ws.segs.n < need) ws.segs.alloc(need);
        if (ws.desc.n < need) ws.desc.alloc(need);
    }
    ZP_LAUNCH(msm_segcount_kernel, dim3((unsigned)((wb + 255) / 256)), dim3(256), 0, st, run_begin, run_end, wb, (uint32_t)ws.seg,
              ws.seg_cnt.p);
    msm_scan(ws.seg_cnt.p, ws.seg_start.p, wb, ws.tile_sum.p, st);
    ZP_LAUNCH(msm_segdesc_kernel, dim3((unsigned)((wb + 255) / 256)), dim3(256), 0, st, run_begin, run_end, ws.seg_start.p, wb,
              (uint32_t)ws.seg, ws.desc.p);
    mark(4);
    {
        ZP_CUDA(cudaMemsetAsync(ws.counter.p, 0, sizeof(uint32_t), st));
        int variant = ws.acc_variant;
        int blocks_per_sm = variant == 4 ? 4 : (variant == 5 ? 5 : 3);
        unsigned grid = (unsigned)(ws.sm_count * blocks_per_sm);
        if (variant == 4) {
            auto k = msm_accumulate_kernel<4>;
            ZP_LAUNCH(k, dim3(grid), dim3(128), 0, st, ps, entries, ws.desc.p, ws.seg_start.p + wb, ws.counter.p, ws.segs.p);
        } else if (variant == 5) {
            auto k = msm_accumulate_kernel<5>;
            ZP_LAUNCH(k, dim3(grid), dim3(128), 0, st, ps, entries, ws.desc.p, ws.seg_start.p + wb, ws.counter.p, ws.segs.p);
        } else {
            auto k = msm_accumulate_kernel<3>;
            ZP_LAUNCH(k, dim3(grid), dim3(128), 0, st, ps, entries, ws.desc.p, ws.seg_start.p + wb, ws.counter.p, ws.segs.p);
        }
    }
    ZP_CUDA(cudaMemsetAsync(ws.counter.p + 1, 0, sizeof(uint32_t), st));
    ZP_LAUNCH(msm_fold_small_kernel, dim3((unsigned)((wb + 127) / 128)), dim3(128), 0, st, ws.segs.p, ws.seg_start.p, wb, ws.fold_list.p,
              ws.counter.p + 1);
    ZP_LAUNCH(msm_fold_kernel, dim3((unsigned)(ws.sm_count * 8)), dim3(128), 0, st, ws.segs.p, ws.seg_start.p, ws.fold_list.p,
              ws.counter.p + 1);
    mark(5);
    const int groups = msm_reduce_groups(cfg), lw2 = msm_reduce_lw2(cfg);
    // threads per row / column sum: fewer threads = longer serial part but a shorter tree and more resident CTAs
    static const int rowcol_threads = getenv("ZP_MSM_ROWCOL_THREADS") ? atoi(getenv("ZP_MSM_ROWCOL_THREADS")) : 32;  // measured 32 / 64 / 128: 5.3 / 5.7 / 7.2 ms for a batch of 6
    ZP_LAUNCH(msm_rowcol_kernel, dim3((unsigned)(nsets * msm_reduce_entries(cfg))), dim3(rowcol_threads), 0, st, ws.segs.p,
              ws.seg_start.p, cfg.nbuckets, lw2, ws.rowcol.p);
    ZP_LAUNCH(msm_weighted_kernel, dim3((unsigned)(nsets * groups)), dim3(128), 0, st, ws.rowcol.p, cfg.nbuckets, lw2, groups,
              ws.partial.p);
    ZP_LAUNCH(msm_final_kernel, dim3(nsets), dim3(128), 0, st, ws.partial.p, groups, ws.final_sums.p);
    if (cfg.bucket_lg) {
        // plain sum of the rank's buckets = sum of its column sums rc[set][W1 .. W1 + W2)
        const int W2 = 1 << lw2, W1 = cfg.nbuckets >> lw2;
        ZP_LAUNCH(msm_plain_total_kernel, dim3(nsets), dim3(128), 0, st, ws.rowcol.p, W1, W2, ws.final_sums.p + nsets);
    }
    mark(6);
    ZP_CUDA(cudaMemcpyAsync(ws.partial_host.data(), ws.final_sums.p, (size_t)(cfg.bucket_lg ? 2 : 1) * nsets * sizeof(xyzz_t),
                            cudaMemcpyDeviceToHost, st));
    if (ws.ba_used) {
        // entries the accumulate kernel saw (for the roofline accounting) + the degenerate-pair flag
        ZP_CUDA(cudaMemcpyAsync(&ws.ba_flag_host[1], run_begin + wb, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        ZP_CUDA(cudaMemcpyAsync(&ws.ba_flag_host[0], ws.ba_flag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        if (ws.timing) ZP_CUDA(cudaMemcpyAsync(&ws.ba_entries_host, ws.start.p + wb, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    ws.acc_entries = (double)wn;
    ws.ba_rounds_used = rounds > 0 ? rounds : 0;
}

void msm_launch(MsmWorkspace& ws, const MsmConfig& cfg, const affine_t* points, const fr_t* scalars, size_t n, cudaStream_t st) {
    MsmBatch b{};
    b.s[0] = scalars;
    msm_launch_impl(ws, cfg, points, b, 1, n, true, st);
}
void msm_launch_batch(MsmWorkspace& ws, const MsmConfig& cfg, const affine_t* points, const fr_t* const* scalars, int nbatch, size_t n,
                      cudaStream_t st) {
    MsmBatch b{};
    for (int k = 0; k < nbatch && k < MSM_MAX_BATCH; k++) b.s[k] = scalars[k];
    msm_launch_impl(ws, cfg, points, b, nbatch, n, true, st);
}

std::vector<host::G1> msm_collect_batch(MsmWorkspace& ws, const MsmConfig& cfg, cudaStream_t st) {
    ZP_CUDA(cudaStreamSynchronize(st));
    if (ws.ba_used) {
        if (ws.ba_flag_host[0]) {
            // a pair with equal x-coordinates: redo this MSM on the exact XYZZ-only path
            msm_launch_impl(ws, cfg, ws.last_points, ws.last_batch, ws.last_nbatch, ws.last_n, false, st);
            ZP_CUDA(cudaStreamSynchronize(st));
        } else {
            ws.acc_entries = (double)ws.ba_flag_host[1];
        }
    }
    if (ws.timing && ws.ev[6]) {
        for (int k = 0; k < 6; k++) {
            float ms = 0;
            ZP_CUDA(cudaEventElapsedTime(&ms, ws.ev[k], ws.ev[k + 1]));
            ws.last_ms[k] = ms;
        }
        ws.down0_ms = ws.down0_pairs = 0;
        ws.down0_launches = 0;
        if (ws.ba_used && ws.ba_rounds_used <= MsmWorkspace::BA_MAX_ROUNDS && ws.down_ev[0]) {
            for (int r = 0; r < ws.ba_rounds_used; r++) {
                float ms = 0;
                ZP_CUDA(cudaEventElapsedTime(&ms, ws.down_ev[2 * r], ws.down_ev[2 * r + 1]));
                ws.down0_ms += ms;
            }
            ws.down0_launches = ws.ba_rounds_used;
            ws.down0_pairs = (double)ws.ba_entries_host - (double)ws.ba_flag_host[1];
        }
    }
    std::vector<host::G1> res(ws.last_nbatch);
    for (int b = 0; b < ws.last_nbatch; b++) {
        const xyzz_t* ph = ws.partial_host.data() + (size_t)b * cfg.nsets;
        host::G1 total = host::G1::infinity();
        if (cfg.nsets == 1) {
            total = host::G1::from_dev(ph[0]);  // precomputed tables: window weights are in the points
            if (cfg.bucket_lg) {
                // cyclic bucket share: true weight of local bucket j is j G + rank + 1  =>  G W - (G - 1 - rank) T
                host::G1 plain = host::G1::from_dev(ws.partial_host[(size_t)ws.last_nbatch + b]);
                for (int k = 0; k < cfg.bucket_lg; k++) total.dbl_inplace();
                const uint32_t m = (1u << cfg.bucket_lg) - 1 - cfg.bucket_rank;
                if (m) {
                    host::G1 acc = host::G1::infinity();
                    for (int bit = 31 - __builtin_clz(m); bit >= 0; bit--) {
                        acc.dbl_inplace();
                        if ((m >> bit) & 1) acc.add(plain);
                    }
                    acc.Y = acc.Y.neg();
                    total.add(acc);
                }
            }
        } else {
            for (int w = cfg.nwin - 1; w >= 0; w--) {
                for (int k = 0; k < cfg.c; k++) total.dbl_inplace();
                total.add(host::G1::from_dev(ph[w]));
            }
        }
        res[b] = total;
    }
    return res;
}
host::G1 msm_collect(MsmWorkspace& ws, const MsmConfig& cfg, cudaStream_t st) { return msm_collect_batch(ws, cfg, st)[0]; }

}  // namespace zp
