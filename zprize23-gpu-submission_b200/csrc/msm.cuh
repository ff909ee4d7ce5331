// G1 multi-scalar multiplication (KZG commitments).  Host-side interface of msm.cu.
// Replaces the reference's `multi_scalar_mult` ("Prize 1B/plonk-core/lib/PLONK/utils/function.cu":275-290
// -> "…/utils/zkp/cuda/zksnark_msm.cu":45-83 -> sppark pippenger.cuh:470-556 -> CPU collect.h:378-445).
#pragma once
#include "common.cuh"
#include "curve.cuh"
#include "host_math.hpp"

namespace zp {

struct MsmConfig {
    int c;              // window bits
    int nwin;           // number of windows = ceil(256 / c)
    int nbuckets;       // buckets per bucket set accumulated by this launch: 2^(c-1), or a power-of-two slice of them
    // multi-GPU sharding by BUCKETS (see msm.cu): this launch owns the global buckets g with g mod 2^bucket_lg == bucket_rank,
    // numbered locally g >> bucket_lg; nbuckets = 2^(c-1) >> bucket_lg.  bucket_lg = 0: all buckets.
    int bucket_lg = 0;
    uint32_t bucket_rank = 0;
    int nsets;          // bucket sets: nwin normally, 1 with precomputed window tables
    size_t tab_stride;  // 0, or the row stride of a precomputed table [w][i] = 2^(c w) * P_i
    uint32_t pt_stride; // bytes between consecutive points: 96 (FFI affine_t) or 128 (padded table entries)
};
// Entry of a precomputed window table: one 128-byte line per point.  DRAM fills L2 in 128-byte lines on B200 (ncu:
// 160 B fetched per random 48-byte x read, 192 B per 96-byte point with the packed 96-byte layout), so the random
// gathers of the bucket accumulation cost exactly one line per point instead of 1.5 on average.
struct alignas(128) affine_pad_t {
    fq_t x, y;
    uint32_t pad[8];
};
MsmConfig msm_config_for(size_t n, int c_override = 0);
// start[0..m] = exclusive scan of cnt[0..m) (cnt is overwritten with the same prefix); tile_sum: m/2048 + 2 words of scratch
void u32_exclusive_scan(uint32_t* cnt, uint32_t* start, size_t m, uint32_t* tile_sum, cudaStream_t st);
// Precomputed-window variant: with T[w][i] = 2^(c w) P_i resident, every window digit of every scalar goes into ONE
// bucket set, so c can grow (fewer windows => fewer bucket additions) without multiplying the bucket count.
MsmConfig msm_config_precomp(size_t n, size_t tab_stride);
// dst[w * n + i] = 2^(c w) * src[i], w < nwin (dst may be larger than 4 GiB; built once per SRS)
void msm_build_table(affine_pad_t* dst, const affine_t* src, size_t n, int c, int nwin, cudaStream_t st);

// Several scalar vectors over the same points go through ONE pipeline (one bucket-set group per member): the
// latency-bound stages (scans, inversion-tree tops, bucket reduction) are paid once per batch instead of once per MSM.
static const int MSM_MAX_BATCH = 8;
struct MsmBatch {
    const fr_t* s[MSM_MAX_BATCH];
};

// a few pinned host words owned by a workspace (top level of the inversion tree, <= 256 nodes: one round trip per
// batch-affine round; [0, 256) coming back, [256, 512) going out)
struct PinnedFq {
    fq_t* p = nullptr;
    PinnedFq() = default;
    PinnedFq(const PinnedFq&) = delete;
    PinnedFq& operator=(const PinnedFq&) = delete;
    ~PinnedFq() {
        if (p) cudaFreeHost(p);
    }
    fq_t* get() {
        if (!p) ZP_CUDA(cudaMallocHost((void**)&p, 512 * sizeof(fq_t)));
        return p;
    }
};

struct MsmWorkspace {
    DevBuf<uint32_t> digits;   // [nwin][n]   |d| | sign << 31
    DevBuf<uint32_t> sorted;   // [<= nwin*n] point index | sign << 31, grouped by (window, bucket)
    DevBuf<uint32_t> start;    // [nwin*nbuckets + 1] exclusive scan of bucket sizes
    DevBuf<uint32_t> cursor;   // [nwin*nbuckets]     running insert position; == bucket end after the scatter
    DevBuf<uint32_t> seg_start;// [nwin*nbuckets + 1] exclusive scan of ceil(bucket size / seg) (work segments)
    DevBuf<uint32_t> seg_cnt;  // [nwin*nbuckets]
    DevBuf<xyzz_t> segs;       // [<= nwin*n/seg + nwin*nbuckets] partial sum of each work segment
    DevBuf<uint2> desc;        // [<= nwin*n/seg + nwin*nbuckets] (first, last+1) index into `sorted` of each work segment
    DevBuf<uint32_t> counter;  // [0] dynamic segment counter of the persistent accumulate kernel, [1] entries of fold_list
    DevBuf<uint32_t> fold_list;// buckets split into more than 8 work segments (folded by one warp each)
    DevBuf<uint32_t> tile_sum; // scratch of the multi-CTA scan
    int sm_count = 0;
    int acc_variant = 3;       // resident CTAs per SM of the accumulate kernel (ZP_ACC_VARIANT=3|4|5)
    // batch-affine pre-reduction (msm_affine.cuh): rounds of pairwise affine additions before the XYZZ accumulation
    int ba_rounds = 0;            // 0 = not read yet, -1 = off, > 0 = rounds (ZP_MSM_BA_ROUNDS, default 3)
    bool ba_rounds_forced = false;// set by the environment: do not adapt the round count to the bucket load
    size_t ba_min_entries = (size_t)1 << 22;
    DevBuf<affine_t> ba_pts[2];   // materialised partial sums (ping-pong)
    DevBuf<fq_t> ba_pre;          // per-slot prefix products of the leaf groups of the inversion tree
    DevBuf<fq_t> ba_den;          // leaf-group products -> inverses, followed by the upper product-tree levels
    DevBuf<uint32_t> ba_src;      // source index of each output slot
    DevBuf<uint32_t> ba_cnt, ba_rs[2];
    DevBuf<uint32_t> ba_flag;     // [0] degenerate pair seen, [1] entries left for the accumulation
    uint32_t ba_flag_host[2] = {0, 0};
    PinnedFq ba_root;             // top level of the inversion tree: products coming back, inverses going out
    bool ba_used = false;
    double acc_entries = 0;       // bucket entries the accumulate kernel of the last launch processed
    // arguments of the last launch (to redo it on the plain path if a degenerate pair was seen)
    const affine_t* last_points = nullptr;
    MsmBatch last_batch{};
    int last_nbatch = 1;
    size_t last_n = 0;
    int ba_rounds_used = 0;    // batch-affine rounds of the last launch
    // bench statistics of the dominant kernel (ba_down0_kernel), filled when `timing` is set: device time of its
    // launches and the affine additions they performed in the last MSM pipeline (every round turns a run of len entries
    // into ceil(len / 2) with floor(len / 2) additions, so additions = entries before the rounds - entries after them)
    static const int BA_MAX_ROUNDS = 8;
    cudaEvent_t down_ev[2 * BA_MAX_ROUNDS] = {};
    uint32_t ba_entries_host = 0;  // bucket entries before the first round
    double down0_ms = 0, down0_pairs = 0;
    int down0_launches = 0;
    size_t seg = 0;            // points per work segment (2x the mean bucket load, >= 32)
    DevBuf<xyzz_t> rowcol;     // [nsets][W1 + W2] row / column sums of the bucket matrix (bucket reduction, step 1)
    DevBuf<xyzz_t> partial;    // [nsets][groups] weighted partial sums (bucket reduction, step 2)
    DevBuf<xyzz_t> final_sums; // [2 * nsets] per-set weighted sums, then per-set plain sums (what returns to the host)
    std::vector<xyzz_t> partial_host;
    // optional per-stage timing (bench only): digits, scan, scatter, batch-affine rounds, accumulate (+ folds), reduce
    bool timing = false;
    cudaEvent_t ev[7] = {0, 0, 0, 0, 0, 0, 0};
    double last_ms[6] = {0, 0, 0, 0, 0, 0};
    void reserve(size_t n, const MsmConfig& cfg, int nbatch = 1);
};

// result = sum_i scalars[i] * points[i]; scalars are Montgomery Fr (as they live in polynomial buffers;
// the canonical conversion the reference does as a separate `to_base` pass is fused into the digit kernel).
// Returns the partial sums per window on the host; msm_finish() folds them into one point.
// When `shard_lo/shard_hi` restrict the point range, the result is that range's partial sum (multi-GPU).
void msm_launch(MsmWorkspace& ws, const MsmConfig& cfg, const affine_t* points, const fr_t* scalars, size_t n,
                cudaStream_t st);
host::G1 msm_collect(MsmWorkspace& ws, const MsmConfig& cfg, cudaStream_t st);
// batch of <= MSM_MAX_BATCH scalar vectors (device pointers, n each) over the same points; one result per member
void msm_launch_batch(MsmWorkspace& ws, const MsmConfig& cfg, const affine_t* points, const fr_t* const* scalars, int nbatch, size_t n,
                      cudaStream_t st);
std::vector<host::G1> msm_collect_batch(MsmWorkspace& ws, const MsmConfig& cfg, cudaStream_t st);

}  // namespace zp
