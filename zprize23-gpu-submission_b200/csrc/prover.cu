// gen_proof protocol driver: the five PLONK rounds of `Prover::prove_with_preprocessed`
// ("Prize 1B/plonk-core/src/proof_system/prover.rs":171-660; PNP's GPU twin
// "Prize 1B/plonk-core/lib/PLONK/src/gen_proof.cuh":10-489) on top of the resident context.
// Transcript schedule = SURVEY Appendix A.  Everything heavy stays in HBM; per proof only the witness
// columns go host->device and 19 points + 26 scalars come back.
#include "prover.cuh"
#include "gates.cuh"
#include <algorithm>
#include <thread>

namespace zp {

std::atomic<unsigned long long> g_launch_count{0};

using host::Fr;
using host::Fq;

// ---- phase timing with CUDA events on the prover's stream ----------------------------------------
struct PhaseTimer {
    struct Span { int cat; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    cudaStream_t st;
    explicit PhaseTimer(cudaStream_t s) : st(s) {}
    ~PhaseTimer() { for (auto e : pool) cudaEventDestroy(e); }
    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t e;
            ZP_CUDA(cudaEventCreate(&e));
            pool.push_back(e);
        }
        return pool[used++];
    }
    void reset() { spans.clear(); used = 0; }
    int begin(int cat) {
        Span s;
        s.cat = cat;
        s.a = get();
        s.b = get();
        ZP_CUDA(cudaEventRecord(s.a, st));
        spans.push_back(s);
        return (int)spans.size() - 1;
    }
    void end(int id) { ZP_CUDA(cudaEventRecord(spans[id].b, st)); }
    void collect(double* out5) {
        ZP_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < 5; i++) out5[i] = 0;
        for (auto& s : spans) {
            float ms = 0;
            ZP_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
            out5[s.cat] += ms;
        }
    }
};
enum { CAT_TOTAL = 0, CAT_NTT = 1, CAT_MSM = 2, CAT_QUOT = 3, CAT_OTHER = 4 };
// spans are recorded into the timer of the proof in flight on THIS context (Prover::timer; null outside prove_resident)
struct Scope {
    PhaseTimer* t;
    int id;
    Scope(PhaseTimer* timer, int cat) : t(timer), id(timer ? timer->begin(cat) : -1) {}
    ~Scope() { if (t && id >= 0) t->end(id); }
};

// ---- context -----------------------------------------------------------------------------------
Prover::Prover(int logn_) : logn(logn_), n((size_t)1 << logn_), n8((size_t)8 << logn_) {
    const char* pc = getenv("ZP_MSM_PRECOMP");
    if (pc && pc[0] == '0') use_precomp = false;
    const char* cc = getenv("ZP_COSET_COPIES");
    if (cc && cc[0] == '0') coset_copies = false;
    const char* sb = getenv("ZP_SHARD_BUCKETS");
    if (sb && sb[0] == '0') shard_buckets = false;
    const char* pm = getenv("ZP_MSM_PRECOMP_MIN_LOG");
    if (pm) precomp_min = (size_t)1 << atoi(pm);
    if (logn < 6 || logn > NTT_LMAX) throw std::runtime_error("zp_prover_create: log_n must be in [6, 26]");
    // proving needs the 8N extended domain (<= 2^26); a larger context serves the SRS / MSM / NTT operator entry points only
    msm_only = logn + 3 > NTT_LMAX;
#ifndef ZP_EMU
    // experiment knob: DRAM -> L2 fill granularity hint for the random 96-byte point gathers of the MSM (32 | 64 | 128)
    if (const char* g = getenv("ZP_L2_FETCH")) {
        ZP_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g)));
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        fprintf(stderr, "[zprize_b200] L2 fetch granularity set to %zu\n", got);
    }
#endif
    {
        int least = 0, greatest = 0;
        ZP_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        ZP_CUDA(cudaStreamCreateWithPriority(&st, cudaStreamDefault, greatest));
        // ZP_NTT_OVERLAP_PRIO: measurement knob — priority of the second stream (default: the lowest)
        const char* pr = getenv("ZP_NTT_OVERLAP_PRIO");
        ZP_CUDA(cudaStreamCreateWithPriority(&st2, cudaStreamNonBlocking, pr ? atoi(pr) : least));
        for (int k = 0; k < 2; k++) {
            ZP_CUDA(cudaEventCreate(&fork_ev[k]));
            ZP_CUDA(cudaEventCreate(&join_ev[k]));
        }
        for (int k = 0; k < 4; k++) ZP_CUDA(cudaEventCreate(&ov_ev[k]));
        const char* ov = getenv("ZP_NTT_OVERLAP");
        ntt_overlap = ov && ov[0] == '1';
    }
    T.init(st);
    PS.init();
    if (msm_only) {
        ZP_CUDA(cudaStreamSynchronize(st));
        return;
    }
    // L_1 on the coset: coefficients are all 1/N (quotient_poly.rs:346-358)
    l1_coset.alloc(n8);
    DevBuf<fr_t> tmp(n);
    fill(tmp.p, T.ninv[logn], n, st);
    ntt_run(T, NS, NTT_COSET_FWD, logn + 3, tmp.p, n, l1_coset.p, st);
    ZP_CUDA(cudaStreamSynchronize(st));
}
Prover::~Prover() {
    for (int b = 0; b < 2; b++) {
        if (pin_buf[b]) cudaFreeHost(pin_buf[b]);
        if (pin_ev[b]) cudaEventDestroy(pin_ev[b]);
    }
    if (st2) {
        cudaStreamSynchronize(st2);
        cudaStreamDestroy(st2);
    }
    for (int k = 0; k < 2; k++) {
        if (fork_ev[k]) cudaEventDestroy(fork_ev[k]);
        if (join_ev[k]) cudaEventDestroy(join_ev[k]);
    }
    for (int k = 0; k < 4; k++)
        if (ov_ev[k]) cudaEventDestroy(ov_ev[k]);
    if (st && own_stream) cudaStreamDestroy(st);
}

// Forks `count` coset NTTs N -> 8N onto the low-priority stream: they start when everything enqueued on `st` so far (the
// coefficients) is done, and signal join_ev[slot]; the quotient round waits for it.  ov_ev[2 slot], ov_ev[2 slot + 1]
// bracket the work on st2 for the phase report.
void Prover::fork_coset_ntts(int slot, const fr_t* const* in, fr_t* const* out, int count) {
    ZP_CUDA(cudaEventRecord(fork_ev[slot], st));
    ZP_CUDA(cudaStreamWaitEvent(st2, fork_ev[slot], 0));
    ZP_CUDA(cudaEventRecord(ov_ev[2 * slot], st2));
    for (int k = 0; k < count; k++) ntt_run(T, NS2, NTT_COSET_FWD, logn + 3, in[k], n, out[k], st2);
    ZP_CUDA(cudaEventRecord(ov_ev[2 * slot + 1], st2));
    ZP_CUDA(cudaEventRecord(join_ev[slot], st2));
}

static const size_t PIN_CHUNK = (size_t)64 << 20;
void Prover::staged_h2d(void* dst_dev, const void* src_host, size_t bytes) {
    if (bytes < ((size_t)8 << 20)) {
        ZP_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
        return;
    }
    for (int b = 0; b < 2; b++)
        if (!pin_buf[b]) {
            ZP_CUDA(cudaMallocHost(&pin_buf[b], PIN_CHUNK));
            ZP_CUDA(cudaEventCreate(&pin_ev[b]));
            ZP_CUDA(cudaEventRecord(pin_ev[b], st));
        }
    // host threads that fill a staging buffer (ZP_H2D_THREADS; default 8 or what the machine has)
    static const unsigned h2d_env = getenv("ZP_H2D_THREADS") ? (unsigned)atoi(getenv("ZP_H2D_THREADS")) : 8u;
    const unsigned nthreads = std::max(1u, std::min(h2d_env, std::max(1u, std::thread::hardware_concurrency())));
    int b = 0;
    for (size_t off = 0; off < bytes; off += PIN_CHUNK, b ^= 1) {
        const size_t len = std::min(PIN_CHUNK, bytes - off);
        ZP_CUDA(cudaEventSynchronize(pin_ev[b]));  // the DMA that last read this staging buffer has finished
        const char* src = (const char*)src_host + off;
        char* dst = (char*)pin_buf[b];
        std::vector<std::thread> pool;
        const size_t part = (len + nthreads - 1) / nthreads;
        for (unsigned t = 1; t < nthreads; t++)
            if (t * part < len) pool.emplace_back([=] { memcpy(dst + t * part, src + t * part, std::min(part, len - t * part)); });
        memcpy(dst, src, std::min(part, len));
        for (auto& th : pool) th.join();
        ZP_CUDA(cudaMemcpyAsync((char*)dst_dev + off, pin_buf[b], len, cudaMemcpyHostToDevice, st));
        ZP_CUDA(cudaEventRecord(pin_ev[b], st));
    }
}
void Prover::set_stream(cudaStream_t s) {
    ZP_CUDA(cudaStreamSynchronize(st));
    if (st && own_stream) cudaStreamDestroy(st);
    st = s;
    own_stream = false;
}

void Prover::ensure_work_buffers(bool lookup) {
    auto need = [](DevBuf<fr_t>& b, size_t cnt) { if (b.n < cnt) b.alloc(cnt); };
    for (int k = 0; k < 4; k++) {
        need(w_ev[k], n);
        need(w_poly[k], n);
        need(w8[k], n8);
    }
    need(qlk_ev, n);
    need(z_poly, n);
    need(z8, n8);
    need(z2_poly, n);
    need(quot, n8);
    need(t_poly, n8);
    need(num, n);
    need(den, n);
    need(lin, n);
    need(comb, n);
    need(wit, n);
    need(wit2, n);
    if (lookup) {
        need(z28, n8);
        need(t_ev, n); need(f_ev, n); need(h1_ev, n); need(h2_ev, n);
        need(table_poly, n); need(f_poly, n); need(h1_poly, n); need(h2_poly, n);
        need(tb8, n8); need(f8, n8); need(h18, n8); need(h28, n8);
    }
}

void Prover::load_srs(const uint64_t* pts, size_t npts) {
    srs_tab.release();
    tab_n = 0;
    if (npts < n) throw std::runtime_error("zp_prover_load_srs: fewer than N points");
    srs.alloc(n);  // only the first N powers are ever used (load.cu:348-351)
    staged_h2d(srs.p, pts, n * sizeof(affine_t));
    ZP_CUDA(cudaStreamSynchronize(st));
}

// ---- insecure SRS generation (test / benchmark harness; the reference calls KZG10::setup in Rust) ----
__global__ void srs_power_table_kernel(fr_t* table, fr_t base, int shift, int cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    store_fr(&table[i], base.pow_u64((uint64_t)i << shift));
}
__global__ void __launch_bounds__(128) srs_kernel(affine_t* out, size_t cnt, const fr_t* pw_lo, const fr_t* pw_hi,
                                                  const affine_t* pow2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    fr_t s = load_fr(&pw_lo[i & 8191]);
    if (i >> 13) s = s * load_fr(&pw_hi[i >> 13]);
    s = s.from_mont();
    xyzz_t acc = xyzz_t::infinity();
    for (int j = 0; j < 255; j++) {
        if ((s.l[j >> 5] >> (j & 31)) & 1) {
            fq_t x = load_fq(&pow2[j].x), y = load_fq(&pow2[j].y);
            acc.add_affine(x, y);
        }
    }
    // tau^i != 0, so acc is finite
    fq_t inv = (acc.ZZ * acc.ZZZ).inverse();
    fq_t zz_inv = inv * acc.ZZZ, zzz_inv = inv * acc.ZZ;
    store_fq(&out[i].x, acc.X * zz_inv);
    store_fq(&out[i].y, acc.Y * zzz_inv);
}
static void g1_generator_host(Fq& x, Fq& y) {
    // standard BLS12-381 G1 generator, canonical big-endian hex -> Montgomery
    static const uint64_t gx[6] = {0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL, 0xa14e3a3f171bac58ULL,
                                   0xc3688c4f9774b905ULL, 0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL};
    static const uint64_t gy[6] = {0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL, 0x00db18cb2c04b3edULL,
                                   0xfcf5e095d5d00af6ULL, 0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL};
    x = Fq::from_canonical(gx);
    y = Fq::from_canonical(gy);
}
void Prover::generate_srs(const fr_t& tau, size_t npts) {
    if (npts < n) throw std::runtime_error("zp_prover_generate_srs: fewer than N points");
    npts = n;
    srs_tab.release();
    tab_n = 0;
    std::vector<affine_t> pow2(255);
    Fq gx, gy;
    g1_generator_host(gx, gy);
    host::G1 p = host::G1::from_affine(gx, gy);
    for (int j = 0; j < 255; j++) {
        Fq ax, ay;
        bool inf;
        p.to_affine(ax, ay, inf);
        pow2[j].x = host::to_dev(ax);
        pow2[j].y = host::to_dev(ay);
        p.dbl_inplace();
    }
    DevBuf<affine_t> d_pow2(255);
    ZP_CUDA(cudaMemcpyAsync(d_pow2.p, pow2.data(), 255 * sizeof(affine_t), cudaMemcpyHostToDevice, st));
    DevBuf<fr_t> lo(8192), hi((npts >> 13) + 1);
    ZP_LAUNCH(srs_power_table_kernel, dim3(32), dim3(256), 0, st, lo.p, tau, 0, 8192);
    int nhi = (int)((npts >> 13) + 1);
    ZP_LAUNCH(srs_power_table_kernel, dim3((nhi + 255) / 256), dim3(256), 0, st, hi.p, tau, 13, nhi);
    srs.alloc(npts);
    ZP_LAUNCH(srs_kernel, dim3((unsigned)((npts + 127) / 128)), dim3(128), 0, st, srs.p, npts, lo.p, hi.p, d_pow2.p);
    ZP_CUDA(cudaStreamSynchronize(st));
}

// ---- prover key --------------------------------------------------------------------------------
static const uint64_t* pk_coeff_ptr(const ProverKeyC& pk, int i) {
    const uint64_t* v[PK_COUNT] = {pk.q_m_coeffs, pk.q_l_coeffs, pk.q_r_coeffs, pk.q_o_coeffs, pk.q_4_coeffs, pk.q_c_coeffs,
                                   pk.q_hl_coeffs, pk.q_hr_coeffs, pk.q_h4_coeffs, pk.q_arith_coeffs, pk.range_selector_coeffs,
                                   pk.logic_selector_coeffs, pk.fixed_group_add_selector_coeffs,
                                   pk.variable_group_add_selector_coeffs, pk.q_lookup_coeffs, pk.left_sigma_coeffs,
                                   pk.right_sigma_coeffs, pk.out_sigma_coeffs, pk.fourth_sigma_coeffs};
    return v[i];
}
static const uint64_t* pk_eval_ptr(const ProverKeyC& pk, int i) {
    const uint64_t* v[PK_COUNT] = {pk.q_m_evals, pk.q_l_evals, pk.q_r_evals, pk.q_o_evals, pk.q_4_evals, pk.q_c_evals,
                                   pk.q_hl_evals, pk.q_hr_evals, pk.q_h4_evals, pk.q_arith_evals, pk.range_selector_evals,
                                   pk.logic_selector_evals, pk.fixed_group_add_selector_evals,
                                   pk.variable_group_add_selector_evals, pk.q_lookup_evals, pk.left_sigma_evals,
                                   pk.right_sigma_evals, pk.out_sigma_evals, pk.fourth_sigma_evals};
    return v[i];
}

void Prover::load_pk(const ProverKeyC& pk, const uint64_t* coeff_len) {
    if (msm_only) throw std::runtime_error("this context is larger than 2^23: SRS / MSM / NTT operators only, no prover key");
    // reference convention when no lengths are given (gen_proof.cuh:61-62,277-278,319-329)
    static const bool unreadable[PK_COUNT] = {true, false, false, false, false, false, false, false, false, false,
                                              true, true, true, true, true, false, false, false, false};
    DevBuf<fr_t> stage(n8);
    for (int i = 0; i < PK_COUNT; i++) {
        // evaluations (8N): upload, drop if identically zero
        staged_h2d(stage.p, pk_eval_ptr(pk, i), n8 * sizeof(fr_t));
        bool ezero = all_zero(PS, stage.p, n8, st);
        if (ezero) {
            evals[i].release();
        } else {
            evals[i].alloc(n8);
            ZP_CUDA(cudaMemcpyAsync(evals[i].p, stage.p, n8 * sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
        }
        // coefficients (N)
        size_t len = coeff_len ? (size_t)coeff_len[i] : (unreadable[i] ? 0 : n);
        if (len > n) throw std::runtime_error("zp_prover_load_pk: coefficient array longer than N");
        if (ezero) {
            coeffs[i].release();
        } else {
            coeffs[i].alloc(n);
            if (!coeff_len && unreadable[i]) {
                // recover the polynomial from its coset evaluations (degree < N, so the top 7N coefficients vanish)
                DevBuf<fr_t> full(n8);
                ntt_run(T, NS, NTT_COSET_INV, logn + 3, evals[i].p, n8, full.p, st);
                ZP_CUDA(cudaMemcpyAsync(coeffs[i].p, full.p, n * sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
                ZP_CUDA(cudaStreamSynchronize(st));
            } else {
                ZP_CUDA(cudaMemsetAsync(coeffs[i].p, 0, n * sizeof(fr_t), st));
                if (len) staged_h2d(coeffs[i].p, pk_coeff_ptr(pk, i), len * sizeof(fr_t));
            }
        }
        ZP_CUDA(cudaStreamSynchronize(st));
    }
    const uint64_t* tb[4] = {pk.table1, pk.table2, pk.table3, pk.table4};
    table_zero = true;
    for (int c = 0; c < 4; c++) {
        table[c].alloc(n);
        ZP_CUDA(cudaMemcpyAsync(table[c].p, tb[c], n * sizeof(fr_t), cudaMemcpyHostToDevice, st));
        if (!all_zero(PS, table[c].p, n, st)) table_zero = false;
    }
    finish_pk();
}

void Prover::preprocess(const uint64_t* const* selector_evals, const uint64_t* const* tables) {
    preprocess_impl(selector_evals, PK_COUNT, nullptr, tables);
}

// sigma evaluations on H built on the device from the flattened wire map (wiring.cu; permutation/mod.rs:101-215)
void Prover::sigma_from_wiring_host(const uint32_t* vars, const uint32_t* cells, size_t m, uint32_t n_vars, fr_t* const sigma_dev[4]) {
    for (size_t i = 0; i < m; i++) {
        if (vars[i] >= n_vars) throw std::runtime_error("wire map: variable id out of range");
        if ((cells[i] >> 2) >= n) throw std::runtime_error("wire map: gate index outside the domain");
    }
    DevBuf<uint32_t> dv(m ? m : 1), dc(m ? m : 1);
    if (m) {
        ZP_CUDA(cudaMemcpyAsync(dv.p, vars, m * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        ZP_CUDA(cudaMemcpyAsync(dc.p, cells, m * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    }
    sigma_from_wiring(WS, dv.p, dc.p, m, n_vars, logn, T, sigma_dev, st);
    ZP_CUDA(cudaStreamSynchronize(st));
}

void Prover::preprocess_wiring(const uint64_t* const* selector_evals15, const uint32_t* vars, const uint32_t* cells, size_t m,
                               uint32_t n_vars, const uint64_t* const* tables) {
    if (msm_only) throw std::runtime_error("this context is larger than 2^23: SRS / MSM / NTT operators only, no prover key");
    DevBuf<fr_t> sig[4];
    fr_t* sp[4];
    for (int k = 0; k < 4; k++) {
        sig[k].alloc(n);
        sp[k] = sig[k].p;
    }
    sigma_from_wiring_host(vars, cells, m, n_vars, sp);
    preprocess_impl(selector_evals15, 15, sp, tables);
}

// selector_evals: n_host host columns (19: selectors + sigmas, or 15: selectors only with the sigmas in sigma_dev)
void Prover::preprocess_impl(const uint64_t* const* selector_evals, int n_host, const fr_t* const* sigma_dev,
                             const uint64_t* const* tables) {
    if (msm_only) throw std::runtime_error("this context is larger than 2^23: SRS / MSM / NTT operators only, no prover key");
    DevBuf<fr_t> stage(n);
    for (int i = 0; i < PK_COUNT; i++) {
        const fr_t* src = nullptr;
        if (i < n_host) {
            if (selector_evals[i]) {
                ZP_CUDA(cudaMemcpyAsync(stage.p, selector_evals[i], n * sizeof(fr_t), cudaMemcpyHostToDevice, st));
                src = stage.p;
            }
        } else {
            src = sigma_dev[i - PK_SIGL];
        }
        if (!src || all_zero(PS, src, n, st)) {
            coeffs[i].release();
            evals[i].release();
            continue;
        }
        coeffs[i].alloc(n);
        evals[i].alloc(n8);
        ntt_run(T, NS, NTT_INV, logn, src, n, coeffs[i].p, st);                      // preprocess.rs:345-405
        ntt_run(T, NS, NTT_COSET_FWD, logn + 3, coeffs[i].p, n, evals[i].p, st);    // preprocess.rs:171-246
        ZP_CUDA(cudaStreamSynchronize(st));
    }
    table_zero = true;
    for (int c = 0; c < 4; c++) {
        table[c].alloc(n);
        if (tables && tables[c]) {
            ZP_CUDA(cudaMemcpyAsync(table[c].p, tables[c], n * sizeof(fr_t), cudaMemcpyHostToDevice, st));
            if (!all_zero(PS, table[c].p, n, st)) table_zero = false;
        } else {
            ZP_CUDA(cudaMemsetAsync(table[c].p, 0, n * sizeof(fr_t), st));
        }
    }
    finish_pk();
}

void Prover::finish_pk() {
    for (int k = 0; k < 4; k++) {
        if (!coeffs[PK_SIGL + k].p) throw std::runtime_error("prover key: a sigma polynomial is identically zero");
        sigma_h[k].alloc(n);
        ntt_run(T, NS, NTT_FWD, logn, coeffs[PK_SIGL + k].p, n, sigma_h[k].p, st);  // permutation/mod.rs:669-674
    }
    ZP_CUDA(cudaStreamSynchronize(st));
    have_pk = true;
}

// sum_{i in [lo, hi)} scalars[i - lo] * srs[i]; large ranges go through the precomputed window table of the
// slice [lo, lo + slice) (built on first use; 13 x 384 MiB at N = 2^22)
host::G1 Prover::msm_over_srs(const fr_t* scalars_dev, size_t lo, size_t hi, size_t slice) {
    return msm_over_srs_batch(&scalars_dev, 1, lo, hi, slice)[0];
}
// true when a sharded commitment of ncoef coefficients splits the BUCKETS of the precomputed-table MSM across the ranks
// (every rank walks all points; see msm.cu) instead of the point range
bool Prover::shard_by_buckets(size_t ncoef) const {
    // below 2^21 coefficients a rank's bucket share is too small for the batch-affine rounds and point ranges win
    // (profiles/r02r_msm_sharded_sweep.jsonl: 2^20 on 8 GPUs 6.6 ms by buckets, 2.3 ms by points; 2^22: 5.3 vs 5.1 single,
    // but 43 vs 49 ms per proof with its 4- and 6-member batches)
    static const int min_log = getenv("ZP_SHARD_BUCKETS_MIN_LOG") ? atoi(getenv("ZP_SHARD_BUCKETS_MIN_LOG")) : 21;
    if (shard_world <= 1 || (shard_world & (shard_world - 1)) || !shard_buckets || !use_precomp || ncoef < precomp_min ||
        ncoef < ((size_t)1 << min_log))
        return false;
    MsmConfig c = msm_config_precomp(srs.n, srs.n);
    return c.nbuckets / shard_world >= 256;
}
// the same for k scalar vectors at once (one MSM pipeline, see MsmBatch); bucket_world > 1: only the bucket slice of
// bucket_rank (precomputed-table route)
std::vector<host::G1> Prover::msm_over_srs_batch(const fr_t* const* scalars_dev, int k, size_t lo, size_t hi, size_t slice,
                                                 int bucket_rank, int bucket_world) {
    MsmConfig cfg = msm_config_for(hi - lo);
    // measurement knob: time one rank's share of a bucket-sharded MSM on a single GPU (tools/bench_msm.py)
    if (bucket_world == 1 && getenv("ZP_BENCH_BUCKET_WORLD")) {
        bucket_world = atoi(getenv("ZP_BENCH_BUCKET_WORLD"));
        bucket_rank = getenv("ZP_BENCH_BUCKET_RANK") ? atoi(getenv("ZP_BENCH_BUCKET_RANK")) : bucket_world - 1;
    }
    const affine_t* base = srs.p + lo;
    if (use_precomp && hi - lo >= precomp_min) {
        if (!(srs_tab.p && tab_lo == lo && tab_n >= hi - lo)) {
            if (slice < hi - lo) slice = hi - lo;
            tab_cfg = msm_config_precomp(slice, slice);
            srs_tab.alloc((size_t)tab_cfg.nwin * slice);
            msm_build_table(srs_tab.p, srs.p + lo, slice, tab_cfg.c, tab_cfg.nwin, st);
            tab_lo = lo;
            tab_n = slice;
        }
        cfg = tab_cfg;
        base = reinterpret_cast<const affine_t*>(srs_tab.p);
        if (bucket_world > 1) {
            cfg.bucket_lg = ilog2((size_t)bucket_world);
            cfg.bucket_rank = (uint32_t)bucket_rank;
            cfg.nbuckets = tab_cfg.nbuckets >> cfg.bucket_lg;
        }
    } else if (bucket_world > 1) {
        throw std::runtime_error("msm: bucket sharding needs the precomputed-table route");
    }
    msm_launch_batch(MW, cfg, base, scalars_dev, k, hi - lo, st);
    std::vector<host::G1> r = msm_collect_batch(MW, cfg, st);
    if (MW.timing) {
        msm_acc_ms += MW.last_ms[3] + MW.last_ms[4];  // bucket accumulation: batch-affine rounds + XYZZ accumulate + folds
        for (int i = 0; i < 6; i++) msm_all_ms += MW.last_ms[i];
        double entries = (double)(hi - lo) * cfg.nwin * k / (bucket_world > 1 ? bucket_world : 1);
        msm_mads += 10.0 * 588.0 * entries;  // 10 * 588 * M * W
        // multiply-adds actually issued: a batch-affine addition costs ~6.2 Fq products (3 for the shared inversion incl.
        // the tree levels above the leaves, 3 for the chord), an XYZZ mixed addition 10
        double left = MW.ba_used ? MW.acc_entries : entries;
        msm_exec_mads += 588.0 * (6.2 * (entries - left) + 10.0 * left);
        msm_launches++;
        msm_count += k;
        msm_down0_ms += MW.down0_ms;
        msm_down0_pairs += MW.down0_pairs;
        msm_down0_launches += MW.down0_launches;
    }
    return r;
}

void Prover::commit(const fr_t* coeffs_dev, size_t ncoef, CommitmentC* out, Fq* ox, Fq* oy, bool* oinf) {
    CommitmentC* outs[1] = {out};
    Fq x, y;
    bool inf;
    commit_batch(&coeffs_dev, 1, ncoef, outs, &x, &y, &inf);
    if (ox) *ox = x;
    if (oy) *oy = y;
    if (oinf) *oinf = inf;
}

// Commitments to k polynomials of ncoef coefficients each (null pointer / ncoef == 0: the identity).  The MSMs of the
// non-trivial members run as ONE batch; when sharded, the k partial sums of a rank travel in one all-gather.
void Prover::commit_batch(const fr_t* const* coeffs_dev, int k, size_t ncoef, CommitmentC* const* outs, Fq* xs, Fq* ys, bool* infs) {
    Scope sc(timer, CAT_MSM);
    if (!srs.p) throw std::runtime_error("commit: no SRS loaded");
    if (k < 1 || k > MSM_MAX_BATCH) throw std::runtime_error("commit_batch: batch size out of range");
    std::vector<host::G1> res(k, host::G1::infinity());
    std::vector<int> live;
    for (int i = 0; i < k; i++)
        if (coeffs_dev[i] && ncoef) live.push_back(i);
    if (!live.empty()) {
        const int m = (int)live.size();
        std::vector<host::G1> part(m, host::G1::infinity());
        if (shard_by_buckets(ncoef)) {
            // bucket-range shard: all points, this rank's slice of the 2^(c-1) buckets (same window size as one GPU)
            std::vector<const fr_t*> sp(m);
            for (int j = 0; j < m; j++) sp[j] = coeffs_dev[live[j]];
            part = msm_over_srs_batch(sp.data(), m, 0, ncoef, srs.n, shard_rank, shard_world);
        } else {
            // point-range shard of this rank (the whole range when world == 1)
            size_t chunk = (ncoef + shard_world - 1) / shard_world;
            size_t lo = std::min(ncoef, (size_t)shard_rank * chunk), hi = std::min(ncoef, lo + chunk);
            if (hi > lo) {
                std::vector<const fr_t*> sp(m);
                for (int j = 0; j < m; j++) sp[j] = coeffs_dev[live[j]] + lo;
                part = msm_over_srs_batch(sp.data(), m, lo, hi, std::min(chunk, srs.n - lo));
            }
        }
        if (shard_world > 1) {
            if (!allgather) throw std::runtime_error("commit: sharded prover without an all-gather callback");
            std::vector<host::G1> all((size_t)shard_world * m);
            if (allgather(allgather_user, part.data(), all.data(), sizeof(host::G1) * m) != 0)
                throw std::runtime_error("commit: all-gather of MSM partial sums failed");
            for (int j = 0; j < m; j++) {
                host::G1 r = host::G1::infinity();
                for (int w = 0; w < shard_world; w++) r.add(all[(size_t)w * m + j]);
                part[j] = r;
            }
        }
        for (int j = 0; j < m; j++) res[live[j]] = part[j];
    }
    for (int i = 0; i < k; i++) {
        res[i].to_affine(xs[i], ys[i], infs[i]);
        memcpy(outs[i]->x, xs[i].v, 48);
        memcpy(outs[i]->y, ys[i].v, 48);
    }
}

void Prover::verifier_key(uint64_t* out23) {
    if (!have_pk) throw std::runtime_error("verifier_key: no prover key");
    CommitmentC* o = reinterpret_cast<CommitmentC*>(out23);
    for (int i = 0; i < PK_COUNT; i++) commit(coeffs[i].p, coeffs[i].p ? n : 0, &o[i]);
    DevBuf<fr_t> tp(n);
    for (int c = 0; c < 4; c++) {
        ntt_run(T, NS, NTT_INV, logn, table[c].p, n, tp.p, st);
        commit(tp.p, n, &o[PK_COUNT + c]);
    }
}

static inline fr_t D(const Fr& a) { return host::to_dev(a); }
static inline Fr H(const fr_t& a) { return host::to_host(a); }
static void put_fr(uint64_t* dst, const Fr& a) { memcpy(dst, a.v, 32); }

void Prover::exchange_blocks(void* base, size_t bytes_per_rank) {
    if (dev_allgather) {
        if (dev_allgather(dev_allgather_user, base, bytes_per_rank) != 0) throw std::runtime_error("device all-gather failed");
        return;
    }
    for (int r = 0; r < shard_world; r++)
        if (dev_bcast(dev_bcast_user, (char*)base + (size_t)r * bytes_per_rank, bytes_per_rank, r) != 0)
            throw std::runtime_error("device broadcast failed");
}

__global__ void __launch_bounds__(256) stride8_gather_kernel(const fr_t* __restrict__ in, fr_t* __restrict__ out, size_t n, int j) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) store_fr(&out[t], load_fr(&in[8 * t + j]));
}
void Prover::ensure_coset_copies() {
    if (evals_coset_rank == shard_rank) return;
    const unsigned grid = (unsigned)((n + 255) / 256);
    for (int i = 0; i < PK_COUNT; i++) {
        evals_coset[i].release();
        if (!evals[i].p) continue;
        evals_coset[i].alloc(n);
        ZP_LAUNCH(stride8_gather_kernel, dim3(grid), dim3(256), 0, st, evals[i].p, evals_coset[i].p, n, shard_rank);
    }
    l1_coset_c.alloc(n);
    ZP_LAUNCH(stride8_gather_kernel, dim3(grid), dim3(256), 0, st, l1_coset.p, l1_coset_c.p, n, shard_rank);
    evals_coset_rank = shard_rank;
}

void Prover::upload_witness(const CircuitC& c) {
    if (c.n > n || c.n == 0) throw std::runtime_error("zp_prover_prove: circuit size does not fit the domain");
    if (c.intended_pi_pos >= n) throw std::runtime_error("zp_prover_prove: intended_pi_pos lies outside the domain");
    const size_t cn = (size_t)c.n;
    // ---- 0. witness upload (gen_proof.cuh:11-17, load.cu:311-345)
    ensure_work_buffers(false);
    const uint64_t* host[5] = {c.w_l, c.w_r, c.w_o, c.w_4, c.q_lookup};
    fr_t* dev[5] = {w_ev[0].p, w_ev[1].p, w_ev[2].p, w_ev[3].p, qlk_ev.p};
    if (shard_world > 1 && dev_bcast) {
        // every rank holds the same host witness: each one sends only its 1/world slice over PCIe and the slices are
        // exchanged over NVLink (8 ranks pulling 506 MB each from host memory at once cost 21 ms at HEIGHT=15)
        const size_t chunk = (cn + shard_world - 1) / shard_world;
        const size_t lo = std::min(cn, (size_t)shard_rank * chunk), hi = std::min(cn, lo + chunk);
        for (int k = 0; k < 5; k++)
            if (hi > lo) ZP_CUDA(cudaMemcpyAsync(dev[k] + lo, host[k] + 4 * lo, (hi - lo) * sizeof(fr_t), cudaMemcpyHostToDevice, st));
        if (dev_allgather) {
            // equal blocks of `chunk` elements (chunk * world <= N; what lies beyond cn is cleared / ignored below)
            for (int k = 0; k < 5; k++) exchange_blocks(dev[k], chunk * sizeof(fr_t));
        } else {
            for (int r = 0; r < shard_world; r++) {
                const size_t rlo = std::min(cn, (size_t)r * chunk), rhi = std::min(cn, rlo + chunk);
                for (int k = 0; k < 5 && rhi > rlo; k++)
                    if (dev_bcast(dev_bcast_user, dev[k] + rlo, (rhi - rlo) * sizeof(fr_t), r) != 0)
                        throw std::runtime_error("device broadcast of a witness slice failed");
            }
        }
    } else {
        for (int k = 0; k < 5; k++) ZP_CUDA(cudaMemcpyAsync(dev[k], host[k], cn * sizeof(fr_t), cudaMemcpyHostToDevice, st));
    }
    for (int k = 0; k < 4; k++)
        if (cn < n) ZP_CUDA(cudaMemsetAsync(w_ev[k].p + cn, 0, (n - cn) * sizeof(fr_t), st));
    wit_lookup_on = !(table_zero && all_zero(PS, qlk_ev.p, cn, st));
    if (wit_lookup_on) ensure_work_buffers(true);
    wit_n = cn;
    memcpy(wit_pi, c.pi, 32);
    wit_pi_pos = c.intended_pi_pos;
}

void Prover::prove(const CircuitC& c, ProofC* out) {
    if (!have_pk) throw std::runtime_error("zp_prover_prove: no prover key loaded");
    if (!srs.p) throw std::runtime_error("zp_prover_prove: no SRS loaded");
    upload_witness(c);
    prove_resident(out);
}

void Prover::prove_resident(ProofC* out) {
    if (!have_pk) throw std::runtime_error("zp_prover_prove: no prover key loaded");
    if (!srs.p) throw std::runtime_error("zp_prover_prove: no SRS loaded");
    if (!wit_n) throw std::runtime_error("zp_prover_prove_resident: no witness uploaded");
    PhaseTimer proof_timer(st);
    timer = &proof_timer;
    struct TimerGuard { Prover* p; ~TimerGuard() { p->timer = nullptr; } } timer_guard{this};
    int total_id = proof_timer.begin(CAT_TOTAL);
    memset(out, 0, sizeof(ProofC));
    const size_t cn = wit_n;
    const bool lookup_on = wit_lookup_on;
    msm_acc_ms = msm_all_ms = msm_mads = msm_exec_mads = 0;
    msm_launches = msm_count = 0;
    msm_down0_ms = msm_down0_pairs = 0;
    msm_down0_launches = 0;
    MW.timing = collect_msm_stats;
    struct TimingOff { MsmWorkspace& w; ~TimingOff() { w.timing = false; } } timing_off{MW};

    const bool overlap = ntt_overlap && shard_world == 1 && !msm_only;

    MerlinTranscript tr(label);
    Fr pi_val = Fr::from_canonical(wit_pi);  // CircuitC.pi is canonical (prover.rs:721-725)
    std::vector<std::pair<uint64_t, Fr>> pis;
    if (!pi_val.is_zero()) pis.push_back({wit_pi_pos, pi_val});  // PublicInputs keeps non-zero values only
    tr.append_public_inputs("pi", pis);

    // ---- 1. witness polynomials (prover.rs:192-228)
    CommitmentC* comm = &out->a_comm;  // 19 consecutive CommitmentC
    static const char* wl[4] = {"w_l", "w_r", "w_o", "w_4"};
    {
        // the four wire commitments do not depend on each other: one MSM batch
        const fr_t* wp[4];
        CommitmentC* wc[4];
        Fq x[4], y[4]; bool inf[4];
        // multi-GPU with a device broadcast: the four independent iNTTs are dealt round-robin and broadcast (128 MiB each
        // over NVLink is cheaper than the transform); otherwise every rank transforms all four
        static const int deal_min_log = getenv("ZP_DEAL_MIN_LOG") ? atoi(getenv("ZP_DEAL_MIN_LOG")) : 16;
        const bool deal = shard_world > 1 && dev_bcast != nullptr && logn >= deal_min_log;
        for (int k = 0; k < 4; k++) {
            Scope s(timer, CAT_NTT);
            if (!deal || k % shard_world == shard_rank) ntt_run(T, NS, NTT_INV, logn, w_ev[k].p, n, w_poly[k].p, st);
        }
        for (int k = 0; k < 4; k++) {
            if (deal) {
                Scope s(timer, CAT_NTT);
                if (dev_bcast(dev_bcast_user, w_poly[k].p, n * sizeof(fr_t), k % shard_world) != 0)
                    throw std::runtime_error("device broadcast of a wire polynomial failed");
            }
            wp[k] = w_poly[k].p;
            wc[k] = &comm[k];
        }
        if (overlap) {
            fr_t* w8p[4] = {w8[0].p, w8[1].p, w8[2].p, w8[3].p};
            fork_coset_ntts(0, wp, w8p, 4);
        }
        commit_batch(wp, 4, n, wc, x, y, inf);
        for (int k = 0; k < 4; k++) tr.append_point(wl[k], x[k], y[k], inf[k]);
    }

    // ---- 2. lookup polynomials (prover.rs:230-329)
    Fr zeta = tr.challenge_scalar("zeta");
    tr.append_scalar("zeta", zeta);
    {
        Fq x, y; bool inf;
        if (lookup_on) {
            { Scope s(timer, CAT_OTHER);
              compress4(t_ev.p, table[0].p, table[1].p, table[2].p, table[3].p, D(zeta), n, st);
              query_f(f_ev.p, w_ev[0].p, w_ev[1].p, w_ev[2].p, w_ev[3].p, qlk_ev.p, cn, t_ev.p, D(zeta), n, st); }
            { Scope s(timer, CAT_NTT);
              ntt_run(T, NS, NTT_INV, logn, t_ev.p, n, table_poly.p, st);
              ntt_run(T, NS, NTT_INV, logn, f_ev.p, n, f_poly.p, st); }
            commit(f_poly.p, n, &comm[5], &x, &y, &inf);
            tr.append_point("f", x, y, inf);
            // h1, h2 = combine_split(t, f) on the device (hash table keyed by the table value, ordered by first occurrence)
            { Scope s(timer, CAT_OTHER);
              if (!combine_split(CS, t_ev.p, f_ev.p, n, h1_ev.p, h2_ev.p, st))
                  throw std::runtime_error("lookup: query element not in table (Error::ElementNotIndexed)"); }
            { Scope s(timer, CAT_NTT);
              ntt_run(T, NS, NTT_INV, logn, h1_ev.p, n, h1_poly.p, st);
              ntt_run(T, NS, NTT_INV, logn, h2_ev.p, n, h2_poly.p, st); }
            commit(h1_poly.p, n, &comm[6], &x, &y, &inf);
            tr.append_point("h1", x, y, inf);
            commit(h2_poly.p, n, &comm[7], &x, &y, &inf);
            tr.append_point("h2", x, y, inf);
        } else {
            // table == 0 and q_lookup == 0: t = f = h1 = h2 = 0, commitments are the identity
            commit(nullptr, 0, &comm[5], &x, &y, &inf);
            tr.append_point("f", x, y, inf);
            commit(nullptr, 0, &comm[6], &x, &y, &inf);
            tr.append_point("h1", x, y, inf);
            commit(nullptr, 0, &comm[7], &x, &y, &inf);
            tr.append_point("h2", x, y, inf);
        }
    }

    // ---- 3. permutation polynomials (prover.rs:331-397)
    Fr beta = tr.challenge_scalar("beta");
    tr.append_scalar("beta", beta);
    Fr gamma = tr.challenge_scalar("gamma");
    tr.append_scalar("gamma", gamma);
    Fr delta = tr.challenge_scalar("delta");
    tr.append_scalar("delta", delta);
    Fr epsilon = tr.challenge_scalar("epsilon");
    tr.append_scalar("epsilon", epsilon);
    if (beta == gamma || beta == delta || beta == epsilon || gamma == delta || gamma == epsilon || delta == epsilon)
        throw std::runtime_error("challenges must be different");  // prover.rs:348-353
    {
        const fr_t* wp[4] = {w_ev[0].p, w_ev[1].p, w_ev[2].p, w_ev[3].p};
        const fr_t* sp[4] = {sigma_h[0].p, sigma_h[1].p, sigma_h[2].p, sigma_h[3].p};
        { Scope s(timer, CAT_OTHER);
          perm_num_den(num.p, den.p, wp, sp, D(beta), D(gamma), logn, T, st);
          ratio_inplace(num.p, den.p, comb.p, n, st);
          exclusive_prefix_product(PS, den.p, num.p, n, st); }
        { Scope s(timer, CAT_NTT); ntt_run(T, NS, NTT_INV, logn, num.p, n, z_poly.p, st); }
        if (overlap) {
            const fr_t* zin[1] = {z_poly.p};
            fr_t* zout[1] = {z8.p};
            fork_coset_ntts(1, zin, zout, 1);
        }
        Fq x, y; bool inf;
        commit(z_poly.p, n, &comm[4], &x, &y, &inf);
        tr.append_point("z", x, y, inf);
    }
    if (lookup_on) {
        { Scope s(timer, CAT_OTHER);
          lookup_num_den(num.p, den.p, f_ev.p, t_ev.p, h1_ev.p, h2_ev.p, D(delta), D(epsilon), n, st);
          ratio_inplace(num.p, den.p, comb.p, n, st);
          exclusive_prefix_product(PS, den.p, num.p, n, st); }
        { Scope s(timer, CAT_NTT); ntt_run(T, NS, NTT_INV, logn, num.p, n, z2_poly.p, st); }
        commit(z2_poly.p, n, &comm[8]);
    } else {
        // every lookup ratio is (1+d) e * e(1+d) / (e(1+d))^2 = 1  =>  z2 = 1 on H, z2(X) = 1
        ZP_CUDA(cudaMemsetAsync(z2_poly.p, 0, n * sizeof(fr_t), st));
        fr_t one = fr_t::one();
        ZP_CUDA(cudaMemcpyAsync(z2_poly.p, &one, sizeof(fr_t), cudaMemcpyHostToDevice, st));
        ZP_CUDA(cudaStreamSynchronize(st));
        commit(z2_poly.p, 1, &comm[8]);
    }
    // z_2 commitment is NOT appended to the transcript (prover.rs:395-397)

    // public-input polynomial (pi.rs:103-116): never materialised — see QuotientArgs::pi_val

    // ---- 4. quotient polynomial (prover.rs:402-489, quotient_poly.rs:34-206)
    Fr alpha = tr.challenge_scalar("alpha");
    tr.append_scalar("alpha", alpha);
    Fr range_sep = tr.challenge_scalar("range separation challenge");
    tr.append_scalar("range seperation challenge", range_sep);
    Fr logic_sep = tr.challenge_scalar("logic separation challenge");
    tr.append_scalar("logic seperation challenge", logic_sep);
    Fr fixed_sep = tr.challenge_scalar("fixed base separation challenge");
    tr.append_scalar("fixed base separation challenge", fixed_sep);
    Fr var_sep = tr.challenge_scalar("variable base separation challenge");
    tr.append_scalar("variable base separation challenge", var_sep);
    Fr lookup_sep = tr.challenge_scalar("lookup separation challenge");
    tr.append_scalar("lookup separation challenge", lookup_sep);

    // JubJub d (ed-on-bls12-381), Montgomery literal — lib/PLONK/src/bls12_381/edwards.cu:20-31
    static const uint64_t JUBJUB_D[4] = {3049539848285517488ULL, 18189135023605205683ULL, 8793554888777148625ULL,
                                         6339087681201251886ULL};
    Fr coeff_d;
    memcpy(coeff_d.v, JUBJUB_D, 32);
    {
        // Multi-GPU (device broadcast hook set, world divides 8): the extended domain g H_8N is the union of the 8 cosets
        // (g w_8N^j) H_N; rank r owns 8 / world of them and does everything of this round on its cosets only — size-N coset
        // NTTs of the inputs, the fused quotient pass, the size-N coset iNTT — then the per-coset coefficient vectors P_j are
        // exchanged (1 GiB in total at N = 2^22) and every rank finishes with the size-8 DFT across cosets (ntt_combine8).
        // Single GPU: the five (ten with lookups) 8N coset NTTs, one quotient pass, one 8N coset iNTT.
        const bool dist = shard_world > 1 && dev_bcast != nullptr && (8 % shard_world) == 0;
        const int cpr = dist ? 8 / shard_world : 0;  // cosets per rank
        struct Job { const fr_t* in; fr_t* out; };
        std::vector<Job> jobs;
        for (int k = 0; k < 4; k++) jobs.push_back({w_poly[k].p, w8[k].p});
        jobs.push_back({z_poly.p, z8.p});
        if (lookup_on) {
            jobs.push_back({z2_poly.p, z28.p});
            jobs.push_back({f_poly.p, f8.p});
            jobs.push_back({table_poly.p, tb8.p});
            jobs.push_back({h1_poly.p, h18.p});
            jobs.push_back({h2_poly.p, h28.p});
        }
        { Scope s(timer, CAT_NTT);
          if (!dist) {
              // jobs 0..4 (wires, z) were forked onto st2 in rounds 1 and 3 when `overlap` is set: join them here
              for (size_t k = overlap ? 5 : 0; k < jobs.size(); k++) ntt_run(T, NS, NTT_COSET_FWD, logn + 3, jobs[k].in, n, jobs[k].out, st);
              if (overlap) {
                  ZP_CUDA(cudaStreamWaitEvent(st, join_ev[0], 0));
                  ZP_CUDA(cudaStreamWaitEvent(st, join_ev[1], 0));
              }
          } else {
              if (cs_tmp.n < n) cs_tmp.alloc(n);
              if (pj8.n < n8) pj8.alloc(n8);
              for (size_t k = 0; k < jobs.size(); k++)
                  for (int c = 0; c < cpr; c++) {  // compact evaluations on coset j at jobs[k].out + c * N
                      const int j = shard_rank * cpr + c;
                      const fr_t* src = jobs[k].in;
                      if (j) {
                          ntt_coset_shift(T, jobs[k].in, cs_tmp.p, n, logn + 3, j, false, st);
                          src = cs_tmp.p;
                      }
                      ntt_run(T, NS, NTT_COSET_FWD, logn, src, n, jobs[k].out + (size_t)c * n, st);
                  }
          }
        }
        QuotientArgs qa;
        qa.logn = logn;
        for (int k = 0; k < 4; k++) qa.w[k] = w8[k].p;
        qa.z = z8.p;
        qa.z2 = lookup_on ? z28.p : nullptr;
        qa.f = f8.p; qa.table = tb8.p; qa.h1 = h18.p; qa.h2 = h28.p;
        qa.pi_count = pis.empty() ? 0 : 1;
        qa.pi_val = D(pi_val);
        qa.pi_shift = (uint32_t)(8 * wit_pi_pos);
        qa.l1 = l1_coset.p;
        for (int i = 0; i < 15; i++) qa.sel[i] = evals[i].p;
        for (int k = 0; k < 4; k++) qa.sigma[k] = evals[PK_SIGL + k].p;
        qa.alpha_sq = D(alpha.sqr());
        qa.alpha = D(alpha); qa.beta = D(beta); qa.gamma = D(gamma); qa.delta = D(delta); qa.epsilon = D(epsilon);
        qa.zeta = D(zeta); qa.range_sep = D(range_sep); qa.logic_sep = D(logic_sep); qa.fixed_sep = D(fixed_sep);
        qa.var_sep = D(var_sep); qa.lookup_sep = D(lookup_sep);
        // Z_H on the coset: g^N * w8^k - 1 (preprocess.rs:498-520), inverted once per residue
        Fr g = H(fr_generator_host());
        Fr gn = g.pow_u64(n), om8 = H(T.omega[3]), p = gn;
        for (int k = 0; k < 8; k++) {
            qa.vh_inv[k] = D((p - Fr::one()).inverse());
            p = p * om8;
        }
        qa.coeff_d = D(coeff_d);
        qa.w_lo = T.w_lo.p;
        qa.w_hi = T.w_hi.p;
        qa.beta_g = D(beta * g);
        qa.out = quot.p;
        qa.i_begin = 0;
        qa.i_count = n8;
        qa.coset_j = -1;
        qa.coset_lc = 0;
        if (!dist) {
            { Scope s(timer, CAT_QUOT); quotient_evals(qa, st); }
            { Scope s(timer, CAT_NTT); ntt_run(T, NS, NTT_COSET_INV, logn + 3, quot.p, n8, t_poly.p, st); }
        } else {
            {   // one fused pass over this rank's cosets (compact arrays in, compact quotient values out)
                QuotientArgs qc = qa;
                qc.i_count = (size_t)cpr * n;
                qc.coset_j = shard_rank * cpr;
                qc.coset_lc = ilog2((size_t)cpr);
                if (cpr == 1 && coset_copies) {
                    // one coset per rank: read the key streams from this rank's compact copies (built once)
                    ensure_coset_copies();
                    for (int i = 0; i < 15; i++) qc.sel[i] = evals_coset[i].p;
                    for (int k = 0; k < 4; k++) qc.sigma[k] = evals_coset[PK_SIGL + k].p;
                    qc.l1 = l1_coset_c.p;
                    qc.compact_key = 1;
                }
                Scope s(timer, CAT_QUOT);
                quotient_evals(qc, st);
            }
            for (int c = 0; c < cpr; c++) {
                const int j = shard_rank * cpr + c;
                Scope s(timer, CAT_NTT);
                // P_j = coefficients of the quotient restricted to coset j: iNTT_N, * 7^-m (coset iNTT), * w_8N^(-j m)
                fr_t* pj = pj8.p + (size_t)j * n;
                ntt_run(T, NS, NTT_COSET_INV, logn, quot.p + (size_t)c * n, n, pj, st);
                if (j) ntt_coset_shift(T, pj, pj, n, logn + 3, j, true, st);
            }
            { Scope s(timer, CAT_NTT);
              exchange_blocks(pj8.p, (size_t)cpr * n * sizeof(fr_t));
              ntt_combine8(T, pj8.p, t_poly.p, logn, st);
            }
        }
    }
    static const char* tl[8] = {"t_1", "t_2", "t_3", "t_4", "t_5", "t_6", "t_7", "t_8"};
    bool t_zero[8];
    {
        // t_1 .. t_8 are independent: one MSM batch over the non-zero ones (t_7 = t_8 = 0 for this circuit family)
        const fr_t* tp[8];
        CommitmentC* tc[8];
        Fq x[8], y[8]; bool inf[8];
        for (int k = 0; k < 8; k++) {
            t_zero[k] = all_zero(PS, t_poly.p + (size_t)k * n, n, st);
            tp[k] = t_zero[k] ? nullptr : t_poly.p + (size_t)k * n;
            tc[k] = &comm[9 + k];
        }
        commit_batch(tp, 8, n, tc, x, y, inf);
        for (int k = 0; k < 8; k++) tr.append_point(tl[k], x[k], y[k], inf[k]);
    }

    // ---- 5. linearisation (prover.rs:491-572, linearisation_poly.rs:164-372)
    Fr z_ch = tr.challenge_scalar("z");
    tr.append_scalar("z", z_ch);
    Fr omega = H(T.omega[logn]);
    Fr zs = z_ch * omega;
    // evaluation plan: (slot in ProofEvaluationsC, polynomial, point)
    enum { E_A = 0, E_B, E_C, E_D, E_LSIG, E_RSIG, E_OSIG, E_PERM, E_QLOOKUP, E_Z2NEXT, E_H1, E_H1NEXT, E_H2, E_F, E_TABLE,
           E_TABLENEXT, E_QARITH, E_QC, E_QL, E_QR, E_QHL, E_QHR, E_QH4, E_ANEXT, E_BNEXT, E_DNEXT, NUM_E };
    Fr ev[NUM_E];
    for (int i = 0; i < NUM_E; i++) ev[i] = Fr::zero();
    {
        struct Plan { int slot; const fr_t* poly; bool shifted; };
        std::vector<Plan> plan = {
            {E_A, w_poly[0].p, false}, {E_B, w_poly[1].p, false}, {E_C, w_poly[2].p, false}, {E_D, w_poly[3].p, false},
            {E_LSIG, coeffs[PK_SIGL].p, false}, {E_RSIG, coeffs[PK_SIGR].p, false}, {E_OSIG, coeffs[PK_SIGO].p, false},
            {E_PERM, z_poly.p, true}, {E_QARITH, coeffs[PK_QARITH].p, false}, {E_QLOOKUP, coeffs[PK_QLOOKUP].p, false},
            {E_QC, coeffs[PK_QC].p, false}, {E_QL, coeffs[PK_QL].p, false}, {E_QR, coeffs[PK_QR].p, false},
            {E_ANEXT, w_poly[0].p, true}, {E_BNEXT, w_poly[1].p, true}, {E_DNEXT, w_poly[3].p, true},
            {E_QHL, coeffs[PK_QHL].p, false}, {E_QHR, coeffs[PK_QHR].p, false}, {E_QH4, coeffs[PK_QH4].p, false}};
        if (lookup_on) {
            plan.push_back({E_Z2NEXT, z2_poly.p, true});
            plan.push_back({E_H1, h1_poly.p, false});
            plan.push_back({E_H1NEXT, h1_poly.p, true});
            plan.push_back({E_H2, h2_poly.p, false});
            plan.push_back({E_F, f_poly.p, false});
            plan.push_back({E_TABLE, table_poly.p, false});
            plan.push_back({E_TABLENEXT, table_poly.p, true});
        } else {
            ev[E_Z2NEXT] = Fr::one();  // z2(X) = 1
        }
        const fr_t* polys[32];
        fr_t points[32], results[32];
        int slots[32], cnt = 0;
        for (auto& pl : plan) {
            if (!pl.poly) continue;  // identically-zero polynomial evaluates to 0
            polys[cnt] = pl.poly;
            points[cnt] = D(pl.shifted ? zs : z_ch);
            slots[cnt] = pl.slot;
            cnt++;
        }
        static const int deal_min_log = getenv("ZP_DEAL_MIN_LOG") ? atoi(getenv("ZP_DEAL_MIN_LOG")) : 16;
        if (shard_world > 1 && allgather && logn >= deal_min_log) {
            // multi-GPU: the evaluations are independent — dealt round-robin, the 32-byte results all-gathered (host callback)
            Scope s(timer, CAT_OTHER);
            const int per = (cnt + shard_world - 1) / shard_world;
            const fr_t* my_polys[32];
            fr_t my_points[32], my_results[32];
            int mine = 0;
            for (int i = shard_rank; i < cnt; i += shard_world) {
                my_polys[mine] = polys[i];
                my_points[mine] = points[i];
                mine++;
            }
            if (mine) evaluate_many(PS, my_polys, my_points, mine, n, my_results, st);
            std::vector<fr_t> send(per), recv((size_t)per * shard_world);
            for (int k = 0; k < per; k++) send[k] = k < mine ? my_results[k] : fr_t::zero();
            if (allgather(allgather_user, send.data(), recv.data(), sizeof(fr_t) * per) != 0)
                throw std::runtime_error("all-gather of the evaluations failed");
            for (int i = 0; i < cnt; i++) results[i] = recv[(size_t)(i % shard_world) * per + i / shard_world];
        } else {
            Scope s(timer, CAT_OTHER);
            evaluate_many(PS, polys, points, cnt, n, results, st);
        }
        for (int i = 0; i < cnt; i++) ev[slots[i]] = H(results[i]);
    }
    Fr vanishing = z_ch.pow_u64(n) - Fr::one();
    Fr z_to_n = vanishing + Fr::one();
    Fr l1_eval = vanishing * (Fr::from_u64(n) * (z_ch - Fr::one())).inverse();  // proof.rs:647-658
    {
        GateVals<Fr> g;
        g.a = ev[E_A]; g.b = ev[E_B]; g.c = ev[E_C]; g.d = ev[E_D];
        g.a_next = ev[E_ANEXT]; g.b_next = ev[E_BNEXT]; g.d_next = ev[E_DNEXT];
        g.q_l = ev[E_QL]; g.q_r = ev[E_QR]; g.q_c = ev[E_QC];
        std::vector<const fr_t*> lp;
        std::vector<fr_t> ls;
        auto term = [&](const fr_t* poly, const Fr& s) {
            if (!poly) return;
            lp.push_back(poly);
            ls.push_back(D(s));
        };
        auto pow5 = [](const Fr& x) { Fr s = x.sqr(); return s.sqr() * x; };
        Fr qa = ev[E_QARITH];
        term(coeffs[PK_QM].p, g.a * g.b * qa);          // arithmetic.rs:83-101
        term(coeffs[PK_QL].p, g.a * qa);
        term(coeffs[PK_QR].p, g.b * qa);
        term(coeffs[PK_QO].p, g.c * qa);
        term(coeffs[PK_Q4].p, g.d * qa);
        term(coeffs[PK_QHL].p, pow5(g.a) * qa);
        term(coeffs[PK_QHR].p, pow5(g.b) * qa);
        term(coeffs[PK_QH4].p, pow5(g.d) * qa);
        term(coeffs[PK_QC].p, qa);
        if (coeffs[PK_RANGE].p) term(coeffs[PK_RANGE].p, range_constraints(range_sep, g));
        if (coeffs[PK_LOGIC].p) term(coeffs[PK_LOGIC].p, logic_constraints(logic_sep, g));
        if (coeffs[PK_FIXED].p) term(coeffs[PK_FIXED].p, fbsm_constraints(fixed_sep, g, coeff_d));
        if (coeffs[PK_VAR].p) term(coeffs[PK_VAR].p, curve_add_constraints(var_sep, g, coeff_d));
        // lookup.rs:155-215
        Fr lsep_sq = lookup_sep.sqr(), lsep_cu = lookup_sep * lsep_sq;
        Fr opd = delta + Fr::one(), eopd = epsilon * opd;
        term(coeffs[PK_QLOOKUP].p, (lc4(g.a, g.b, g.c, g.d, zeta) - ev[E_F]) * lookup_sep);
        {
            Fr b0 = epsilon + ev[E_F];
            Fr b1 = eopd + ev[E_TABLE] + delta * ev[E_TABLENEXT];
            Fr b2 = l1_eval * lsep_cu;
            term(z2_poly.p, opd * b0 * b1 * lsep_sq + b2);
            Fr c0 = (Fr::zero() - ev[E_Z2NEXT]) * lsep_sq;
            Fr c1 = eopd + ev[E_H2] + delta * ev[E_H1NEXT];
            if (lookup_on) term(h1_poly.p, c0 * c1);
        }
        // permutation.rs:156-296
        {
            Fr beta_z = beta * z_ch;
            Fr k1 = Fr::from_u64(7), k2 = Fr::from_u64(13), k3 = Fr::from_u64(17);
            Fr a0 = g.a + beta_z + gamma;
            Fr a1 = g.b + k1 * beta_z + gamma;
            Fr a2 = g.c + k2 * beta_z + gamma;
            Fr a3 = g.d + k3 * beta_z + gamma;
            term(z_poly.p, a0 * a1 * a2 * a3 * alpha);
            Fr b0 = g.a + beta * ev[E_LSIG] + gamma;
            Fr b1 = g.b + beta * ev[E_RSIG] + gamma;
            Fr b2 = g.c + beta * ev[E_OSIG] + gamma;
            Fr b = b0 * b1 * b2 * (beta * ev[E_PERM]) * alpha;
            term(coeffs[PK_SIG4].p, Fr::zero() - b);
            term(z_poly.p, l1_eval * alpha.sqr());
        }
        // - Z_H(z) * sum_k z^{kN} t_{k+1}(X)
        Fr zp = Fr::one();
        for (int k = 0; k < 8; k++) {
            if (!t_zero[k]) term(t_poly.p + (size_t)k * n, Fr::zero() - (zp * vanishing));
            zp = zp * z_to_n;
        }
        { Scope s(timer, CAT_OTHER); lincomb(lin.p, lp.data(), ls.data(), (int)lp.size(), n, st); }
    }
    // evaluations into the transcript (prover.rs:532-572) and the proof
    static const char* en[NUM_E] = {"a_eval", "b_eval", "c_eval", "d_eval", "left_sig_eval", "right_sig_eval", "out_sig_eval",
                                    "perm_eval", "q_lookup_eval", "lookup_perm_eval", "h_1_eval", "h_1_next_eval", "h_2_eval",
                                    "f_eval", "", "", "q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "q_hl_eval", "q_hr_eval",
                                    "q_h4_eval", "a_next_eval", "b_next_eval", "d_next_eval"};
    static const int order[] = {E_A, E_B, E_C, E_D, E_LSIG, E_RSIG, E_OSIG, E_PERM, E_F, E_QLOOKUP, E_Z2NEXT, E_H1, E_H1NEXT, E_H2,
                                E_QARITH, E_QC, E_QL, E_QR, E_QHL, E_QHR, E_QH4, E_ANEXT, E_BNEXT, E_DNEXT};
    for (int idx : order) tr.append_scalar(en[idx], ev[idx]);
    uint64_t* eout = reinterpret_cast<uint64_t*>(&out->evaluations);
    for (int i = 0; i < NUM_E; i++) put_fr(eout + 4 * i, ev[i]);

    // ---- 6. openings (prover.rs:574-636; kzg10.cu:87-146)
    // both witness polynomials are built first (nothing is appended to the transcript between the two challenges),
    // then committed as one MSM batch
    auto open = [&](const std::vector<const fr_t*>& polys, const Fr& point, const Fr& chal, fr_t* wit_out) {
        std::vector<const fr_t*> lp;
        std::vector<fr_t> ls;
        Fr cj = Fr::one();
        for (const fr_t* p : polys) {
            if (p) {
                lp.push_back(p);
                ls.push_back(D(cj));
            }
            cj = cj * chal;
        }
        { Scope s(timer, CAT_OTHER);
          lincomb(comb.p, lp.data(), ls.data(), (int)lp.size(), n, st);
          divide_by_linear(PS, comb.p, n, D(point), wit_out, st); }
    };
    // multi-GPU with a device broadcast: the two witness polynomials are built by ranks 0 and 1 and broadcast
    static const int open_deal_min_log = getenv("ZP_DEAL_MIN_LOG") ? atoi(getenv("ZP_DEAL_MIN_LOG")) : 16;
    const bool deal_open = shard_world > 1 && dev_bcast != nullptr && logn >= open_deal_min_log;
    Fr aw = tr.challenge_scalar("aggregate_witness");
    if (!deal_open || shard_rank == 0)
        open({lin.p, coeffs[PK_SIGL].p, coeffs[PK_SIGR].p, coeffs[PK_SIGO].p, lookup_on ? f_poly.p : nullptr,
              lookup_on ? h2_poly.p : nullptr, lookup_on ? table_poly.p : nullptr, w_poly[0].p, w_poly[1].p, w_poly[2].p, w_poly[3].p},
             z_ch, aw, wit.p);
    Fr saw = tr.challenge_scalar("aggregate_witness");
    if (!deal_open || shard_rank == 1)
        open({z_poly.p, w_poly[0].p, w_poly[1].p, w_poly[3].p, lookup_on ? h1_poly.p : nullptr, z2_poly.p,
              lookup_on ? table_poly.p : nullptr},
             zs, saw, wit2.p);
    if (deal_open) {
        Scope s(timer, CAT_OTHER);
        if (dev_bcast(dev_bcast_user, wit.p, n * sizeof(fr_t), 0) != 0 || dev_bcast(dev_bcast_user, wit2.p, n * sizeof(fr_t), 1) != 0)
            throw std::runtime_error("device broadcast of an opening witness polynomial failed");
    }
    {
        const fr_t* op[2] = {wit.p, wit2.p};
        CommitmentC* oc[2] = {&out->aw_opening, &out->saw_opening};
        Fq x[2], y[2]; bool inf[2];
        commit_batch(op, 2, n, oc, x, y, inf);
    }

    proof_timer.end(total_id);
    proof_timer.collect(last_ms);
    last_ms[CAT_OTHER] = last_ms[CAT_TOTAL] - last_ms[CAT_NTT] - last_ms[CAT_MSM] - last_ms[CAT_QUOT];
    last_ms[5] = 0;
    if (overlap) {  // st has waited for both join events and collect() synchronised st: the st2 events are complete
        for (int k = 0; k < 2; k++) {
            float ms = 0;
            ZP_CUDA(cudaEventElapsedTime(&ms, ov_ev[2 * k], ov_ev[2 * k + 1]));
            last_ms[5] += ms;
        }
    }
}

}  // namespace zp
