// Resident prover context and the gen_proof protocol driver (interface of prover.cu).
#pragma once
#include "../../include/zprize_b200.h"
#include "common.cuh"
#include "ntt.cuh"
#include "msm.cuh"
#include "poly.cuh"
#include "wiring.cuh"
#include "transcript.hpp"
#include <string>

namespace zp {

// ProverKeyC polynomial order: q_m,q_l,q_r,q_o,q_4,q_c,q_hl,q_hr,q_h4,q_arith,range,logic,fixed,variable,
// q_lookup, left/right/out/fourth sigma
enum PkPoly {
    PK_QM = 0, PK_QL, PK_QR, PK_QO, PK_Q4, PK_QC, PK_QHL, PK_QHR, PK_QH4, PK_QARITH, PK_RANGE, PK_LOGIC, PK_FIXED, PK_VAR,
    PK_QLOOKUP, PK_SIGL, PK_SIGR, PK_SIGO, PK_SIG4, PK_COUNT
};

struct PhaseTimer;

struct Prover {
    int logn;
    size_t n, n8;
    cudaStream_t st = 0;
    bool own_stream = true;
    std::string label = "Merkle tree";
    // EXPERIMENT, off by default (ZP_NTT_OVERLAP=1 enables; parity-green, measured 0.9 % SLOWER at HEIGHT=15,
    // profiles/r02w_ntt_overlap_second_stream.log).  Second stream, lowest priority (the prover's own stream is created with
    // the highest): the extended-domain coset NTTs of the wire polynomials and of z(X) need no challenge, so they are forked
    // as soon as the coefficients exist, run concurrently with the commitment MSMs and are joined by events before the
    // quotient pass.  Single GPU only (the sharded quotient round transforms per coset).  Why it does not pay: the MSM kernels
    // fill the SMs (registers / thread slots) even while they wait on DRAM, and both workloads are bound by the same integer
    // multiplier — the NTT CTAs only get the slots the MSM gives up, and the MSM phase grows by what the NTT phase shrinks.
    cudaStream_t st2 = 0;
    cudaEvent_t fork_ev[2] = {nullptr, nullptr}, join_ev[2] = {nullptr, nullptr}, ov_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ntt_overlap = false;
    NttScratch NS2;                  // ping-pong scratch of the transforms on st2
    void fork_coset_ntts(int slot, const fr_t* const* in, fr_t* const* out, int count);
    NttTables T;
    NttScratch NS;
    PolyScratch PS;
    CombineSplitScratch CS;
    WiringScratch WS;
    MsmWorkspace MW;

    // ---- resident inputs (uploaded / built once, reused by every proof)
    DevBuf<affine_t> srs;
    // precomputed window table of this rank's SRS slice [tab_lo, tab_lo + tab_n): T[w][i] = 2^(c w) * srs[tab_lo + i]
    DevBuf<affine_pad_t> srs_tab;
    size_t tab_lo = 0, tab_n = 0;
    MsmConfig tab_cfg;
    bool use_precomp = true;
    size_t precomp_min = (size_t)1 << 16;  // smallest range that uses the table (ZP_MSM_PRECOMP_MIN_LOG)
    DevBuf<fr_t> coeffs[PK_COUNT];   // N each; empty = identically zero
    DevBuf<fr_t> evals[PK_COUNT];    // 8N each; empty = identically zero
    DevBuf<fr_t> sigma_h[4];         // sigma evaluations on H
    DevBuf<fr_t> table[4];           // padded lookup columns on H
    bool table_zero = true;          // all four columns identically zero
    DevBuf<fr_t> l1_coset;           // L_1 on g*H_8N
    bool have_pk = false;
    bool msm_only = false;           // log_n in (23, 26]: operator entry points only (the 8N domain would exceed 2^26)

    // ---- per-proof work buffers (allocated once)
    DevBuf<fr_t> w_ev[4], w_poly[4], w8[4];
    DevBuf<fr_t> qlk_ev;
    DevBuf<fr_t> z_poly, z8, z2_poly, z28;
    DevBuf<fr_t> t_ev, f_ev, h1_ev, h2_ev, table_poly, f_poly, h1_poly, h2_poly, tb8, f8, h18, h28;
    DevBuf<fr_t> quot, t_poly;
    DevBuf<fr_t> cs_tmp, pj8;        // multi-GPU quotient round: shifted coefficients (N), per-coset quotient coefficients (8N)
    DevBuf<fr_t> num, den, lin, comb, wit, wit2;
    double last_ms[6] = {0, 0, 0, 0, 0, 0};  // total, NTT, MSM, quotient, other (own stream); [5] NTT work on st2 (overlapped)
    PhaseTimer* timer = nullptr;     // phase timer of the proof in flight on this context
    // witness currently resident in w_ev / qlk_ev (set by upload_witness)
    size_t wit_n = 0;
    uint64_t wit_pi[4] = {0, 0, 0, 0};
    uint64_t wit_pi_pos = 0;
    bool wit_lookup_on = false;
    // per-proof MSM statistics (bucket-accumulation kernel, the dominant kernel of gen_proof)
    bool collect_msm_stats = false;
    double msm_acc_ms = 0, msm_all_ms = 0, msm_mads = 0, msm_exec_mads = 0;
    int msm_launches = 0, msm_count = 0;  // MSM pipelines launched / commitments they produced
    double msm_down0_ms = 0, msm_down0_pairs = 0;  // ba_down0_kernel: device ms and affine additions over the proof
    int msm_down0_launches = 0;

    // multi-GPU: every rank runs the whole protocol on identical inputs, but each KZG commitment's MSM is
    // sharded by point range; the per-rank partial sums (one XYZZ point, 192 B) are exchanged through the
    // caller-supplied all-gather (torch.distributed / NCCL in bench.py) and folded in rank order.
    int shard_rank = 0, shard_world = 1;
    bool shard_buckets = true;  // precomputed-table MSMs: split the bucket range across ranks, not the points (ZP_SHARD_BUCKETS=0: points)
    zp_allgather_fn allgather = nullptr;
    void* allgather_user = nullptr;
    zp_dev_broadcast_fn dev_bcast = nullptr;
    void* dev_bcast_user = nullptr;
    zp_dev_allgather_fn dev_allgather = nullptr;  // in-place all-gather of device memory (one collective instead of G broadcasts)
    void* dev_allgather_user = nullptr;
    // rank r's block [r * bytes, (r + 1) * bytes) of `base` goes to every rank: all-gather hook, else G broadcasts
    void exchange_blocks(void* base, size_t bytes_per_rank);
    // multi-GPU quotient round with one coset per rank (world = 8): compact copies sel_c[i][t] = evals[i][8 t + rank] of the
    // prover-key streams, so the fused pass reads N contiguous elements per stream instead of every 8th of 8N (4x DRAM
    // over-fetch through 128-byte lines)
    DevBuf<fr_t> evals_coset[PK_COUNT];
    DevBuf<fr_t> l1_coset_c;
    int evals_coset_rank = -1;
    bool coset_copies = true;  // ZP_COSET_COPIES=0: read the natural-order streams at stride 8 instead
    void ensure_coset_copies();

    // host (pageable) -> device copies of the big key arrays go through two pinned staging buffers: a few host threads fill
    // one while the DMA engine drains the other (a plain cudaMemcpy from pageable memory stages single-threaded)
    void* pin_buf[2] = {nullptr, nullptr};
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    void staged_h2d(void* dst_dev, const void* src_host, size_t bytes);

    explicit Prover(int logn_);
    ~Prover();
    void set_stream(cudaStream_t s);
    void ensure_work_buffers(bool lookup);
    void load_srs(const uint64_t* pts, size_t npts);
    void generate_srs(const fr_t& tau, size_t npts);
    void load_pk(const ProverKeyC& pk, const uint64_t* coeff_len);
    void preprocess(const uint64_t* const* selector_evals, const uint64_t* const* tables);
    void preprocess_wiring(const uint64_t* const* selector_evals15, const uint32_t* vars, const uint32_t* cells, size_t m,
                           uint32_t n_vars, const uint64_t* const* tables);
    void preprocess_impl(const uint64_t* const* selector_evals, int n_host, const fr_t* const* sigma_dev, const uint64_t* const* tables);
    void sigma_from_wiring_host(const uint32_t* vars, const uint32_t* cells, size_t m, uint32_t n_vars, fr_t* const sigma_dev[4]);
    void finish_pk();
    void verifier_key(uint64_t* out23);
    void upload_witness(const CircuitC& c);
    void synthesize_merkle_witness(int height, const uint64_t* leaves, const uint64_t* params, const uint64_t* blinding,
                                   uint64_t* root_out);
    void read_witness(int k, uint64_t* out);
    void prove_resident(ProofC* out);
    void prove(const CircuitC& c, ProofC* out);

    host::G1 msm_over_srs(const fr_t* scalars_dev, size_t lo, size_t hi, size_t slice);
    std::vector<host::G1> msm_over_srs_batch(const fr_t* const* scalars_dev, int k, size_t lo, size_t hi, size_t slice,
                                             int bucket_rank = 0, int bucket_world = 1);
    bool shard_by_buckets(size_t ncoef) const;
    void commit_batch(const fr_t* const* coeffs_dev, int k, size_t ncoef, CommitmentC* const* outs, host::Fq* xs, host::Fq* ys, bool* infs);
    // commit to n coefficients (Montgomery) on the device; returns affine point (host)
    void commit(const fr_t* coeffs_dev, size_t ncoef, CommitmentC* out, host::Fq* ox = nullptr, host::Fq* oy = nullptr, bool* oinf = nullptr);
};

}  // namespace zp
