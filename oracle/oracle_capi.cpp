// ORACLE — TEST INFRASTRUCTURE ONLY (see zp_field.hpp header).
// C API over the CPU restatement so that tests/, smoke() and bench.py's cpu_baseline leg can drive it
// through ctypes.  Never linked into the product library.
#include "zp_verifier.hpp"
#include <chrono>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace zpo;

namespace {

struct OracleCtx {
    Composer cs;
    ProverKeyO pk;
    std::vector<Fr> sel_evals[NUM_PK_POLYS];  // padded to N, on H (inputs of preprocessing)
    std::vector<G1Affine> srs;
    std::vector<uint64_t> srs_raw;  // 12 u64 per point (x || y, Montgomery) = CommitKeyC.powers_of_g layout
    Fr tau;
    std::vector<Fr> w[4];
    uint64_t pi_canonical[4];
    uint64_t pi_pos;
    VerifierKeyO vk;
    bool have_vk = false, have_pk = false;
    std::string label = "Merkle tree";
};

void proof_to_bytes(const ProofO& p, uint64_t* out) {
    for (int c = 0; c < NUM_COMM; c++) {
        memcpy(out + 12 * c, p.comm[c].x.v, 48);
        memcpy(out + 12 * c + 6, p.comm[c].y.v, 48);
    }
    for (int e = 0; e < NUM_EVAL; e++) memcpy(out + 12 * NUM_COMM + 4 * e, p.eval[e].v, 32);
}
ProofO proof_from_bytes(const uint64_t* in) {
    ProofO p;
    for (int c = 0; c < NUM_COMM; c++) {
        memcpy(p.comm[c].x.v, in + 12 * c, 48);
        memcpy(p.comm[c].y.v, in + 12 * c + 6, 48);
        // FFI encoding of infinity: (0, Mont(1)) — not a curve point, so unambiguous
        p.comm[c].inf = p.comm[c].x.is_zero() && p.comm[c].y == Fq::one();
    }
    for (int e = 0; e < NUM_EVAL; e++) memcpy(p.eval[e].v, in + 12 * NUM_COMM + 4 * e, 32);
    return p;
}

}  // namespace

extern "C" {

int zpo_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// blst acceleration of the CPU baseline (zp_field.hpp): returns 1 when active after the call
int zpo_set_blst(int on) {
    ensure_init();
    blst_api().on = on && blst_api().loaded;
    return blst_api().on ? 1 : 0;
}
int zpo_blst_active() {
    ensure_init();
    return blst_api().on ? 1 : 0;
}
void zpo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// ---- constants (for pinning against the reference's literals) -------------------------------
void zpo_fr_constants(uint64_t* modulus, uint64_t* one, uint64_t* rr, uint64_t* inv, uint64_t* two_adic_root,
                      uint64_t* generator) {
    ensure_init();
    memcpy(modulus, Fr::P().p, 32);
    memcpy(one, Fr::P().one, 32);
    memcpy(rr, Fr::P().rr, 32);
    *inv = Fr::P().inv;
    memcpy(two_adic_root, fr_two_adic_root().v, 32);
    memcpy(generator, fr_generator().v, 32);
}
void zpo_fq_constants(uint64_t* modulus, uint64_t* one, uint64_t* rr, uint64_t* inv) {
    ensure_init();
    memcpy(modulus, Fq::P().p, 48);
    memcpy(one, Fq::P().one, 48);
    memcpy(rr, Fq::P().rr, 48);
    *inv = Fq::P().inv;
}
void zpo_fr_root_of_unity(int logn, uint64_t* fwd, uint64_t* inv, uint64_t* n_inv) {
    ensure_init();
    Fr w = fr_root_of_unity(logn);
    memcpy(fwd, w.v, 32);
    memcpy(inv, w.inverse().v, 32);
    memcpy(n_inv, Fr::from_u64((uint64_t)1 << logn).inverse().v, 32);
}
void zpo_jubjub(uint64_t* a, uint64_t* d) {
    ensure_init();
    memcpy(a, jubjub_a().v, 32);
    memcpy(d, jubjub_d().v, 32);
}
void zpo_g1_generator(uint64_t* xy) {
    ensure_init();
    G1Affine g = g1_generator();
    memcpy(xy, g.x.v, 48);
    memcpy(xy + 6, g.y.v, 48);
}

// ---- element-wise field ops on Montgomery arrays -------------------------------------------
// op: 0 add, 1 sub, 2 mul, 3 inverse(a), 4 to_canonical(a), 5 from_canonical(a), 6 neg(a), 7 a^5
void zpo_fr_op(int op, size_t n, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    ensure_init();
    for (size_t i = 0; i < n; i++) {
        Fr x, y, r;
        memcpy(x.v, a + 4 * i, 32);
        if (b) memcpy(y.v, b + 4 * i, 32);
        switch (op) {
            case 0: r = x + y; break;
            case 1: r = x - y; break;
            case 2: r = x * y; break;
            case 3: r = x.inverse(); break;
            case 4: x.to_canonical(r.v); break;
            case 5: r = Fr::from_canonical(x.v); break;
            case 6: r = -x; break;
            default: r = x.pow_u64(5); break;
        }
        memcpy(out + 4 * i, r.v, 32);
    }
}
void zpo_fq_op(int op, size_t n, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    ensure_init();
    for (size_t i = 0; i < n; i++) {
        Fq x, y, r;
        memcpy(x.v, a + 6 * i, 48);
        if (b) memcpy(y.v, b + 6 * i, 48);
        switch (op) {
            case 0: r = x + y; break;
            case 1: r = x - y; break;
            case 2: r = x * y; break;
            case 3: r = x.inverse(); break;
            case 4: x.to_canonical(r.v); break;
            case 5: r = Fq::from_canonical(x.v); break;
            default: r = -x; break;
        }
        memcpy(out + 6 * i, r.v, 48);
    }
}
// seeded uniform Fr elements (Montgomery) — SURVEY §8d "seed 1 / seed 2" inputs
void zpo_random_fr(uint64_t seed, size_t n, uint64_t* out) {
    ensure_init();
    const size_t CH = 1 << 14;
    size_t nch = (n + CH - 1) / CH;
#pragma omp parallel for schedule(static)
    for (long k = 0; k < (long)nch; k++) {
        SplitMix64 rng(seed * 0x100000001b3ULL + (uint64_t)k);
        for (size_t i = k * CH; i < std::min(n, (size_t)(k + 1) * CH); i++) memcpy(out + 4 * i, rng.next_fr().v, 32);
    }
}

// ---- NTT family: kind 0 fft, 1 ifft, 2 coset_fft, 3 coset_ifft; in-place on n = 2^logn elements ----
void zpo_ntt(int kind, int logn, uint64_t* data) {
    Domain dom(logn);
    std::vector<Fr> a(dom.n);
    memcpy(a.data(), data, 32 * dom.n);
    std::vector<Fr> r = kind == 0 ? dom.fft(a) : kind == 1 ? dom.ifft(a) : kind == 2 ? dom.coset_fft(a) : dom.coset_ifft(a);
    memcpy(data, r.data(), 32 * dom.n);
}
// Horner evaluation of a coefficient array at a point
void zpo_poly_eval(size_t n, const uint64_t* coeffs, const uint64_t* point, uint64_t* out) {
    ensure_init();
    std::vector<Fr> c(n);
    memcpy(c.data(), coeffs, 32 * n);
    Fr z;
    memcpy(z.v, point, 32);
    Fr r = poly_eval(c, z);
    memcpy(out, r.v, 32);
}

// ---- G1 ---------------------------------------------------------------------------------------
// points: n * 12 u64 (x||y Montgomery, no infinity flag); scalars: n * 4 u64 Montgomery Fr.
// out: 12 u64 affine Montgomery, infinity = (0, Mont(1)).  Returns 1 if the result is infinity.
int zpo_msm(size_t n, const uint64_t* points, const uint64_t* scalars, uint64_t* out) {
    ensure_init();
    std::vector<G1Affine> p(n);
    std::vector<Fr> s(n);
    for (size_t i = 0; i < n; i++) {
        memcpy(p[i].x.v, points + 12 * i, 48);
        memcpy(p[i].y.v, points + 12 * i + 6, 48);
        p[i].inf = false;
        memcpy(s[i].v, scalars + 4 * i, 32);
    }
    G1Affine r = g1_msm(p.data(), s.data(), n).to_affine();
    memcpy(out, r.x.v, 48);
    memcpy(out + 6, r.y.v, 48);
    return r.inf ? 1 : 0;
}
// powers_of_g[i] = tau^i * G, tau derived from the seed (SplitMix64 -> Fr)
void zpo_srs(uint64_t tau_seed, size_t n, uint64_t* out_points, uint64_t* out_tau) {
    ensure_init();
    SplitMix64 rng(tau_seed);
    Fr tau = rng.next_fr();
    std::vector<G1Affine> srs = srs_from_tau(tau, n);
    for (size_t i = 0; i < n; i++) {
        memcpy(out_points + 12 * i, srs[i].x.v, 48);
        memcpy(out_points + 12 * i + 6, srs[i].y.v, 48);
    }
    if (out_tau) memcpy(out_tau, tau.v, 32);
}
// scalar * point (affine in, affine out)
void zpo_g1_mul(const uint64_t* point, const uint64_t* scalar, uint64_t* out) {
    ensure_init();
    G1Affine p;
    memcpy(p.x.v, point, 48);
    memcpy(p.y.v, point + 6, 48);
    p.inf = p.x.is_zero() && p.y == Fq::one();
    Fr s;
    memcpy(s.v, scalar, 32);
    G1Affine r = G1::from_affine(p).mul(s).to_affine();
    memcpy(out, r.x.v, 48);
    memcpy(out + 6, r.y.v, 48);
}
int zpo_g1_on_curve(const uint64_t* point) {
    ensure_init();
    G1Affine p;
    memcpy(p.x.v, point, 48);
    memcpy(p.y.v, point + 6, 48);
    p.inf = false;
    return g1_on_curve(p) ? 1 : 0;
}
void zpo_g1_serialize(const uint64_t* point, uint8_t* out48) {
    ensure_init();
    G1Affine p;
    memcpy(p.x.v, point, 48);
    memcpy(p.y.v, point + 6, 48);
    p.inf = p.x.is_zero() && p.y == Fq::one();
    serialize_g1(p, out48);
}

// ---- transcript --------------------------------------------------------------------------------
// Replays: Transcript::new(proto); append_message(label, data); challenge_bytes(chal_label, out[n])
void zpo_transcript_kat(const char* proto, const char* label, const uint8_t* data, size_t data_len, const char* chal_label,
                        uint8_t* out, size_t out_len) {
    Transcript tr(proto);
    tr.append_message(label, data, data_len);
    tr.challenge_bytes(chal_label, out, out_len);
}
// Generic script: ops encoded as (kind u8: 0 append, 1 challenge)(label_len u32)(label)(len u32)(payload if append);
// challenge outputs are concatenated into out.
void zpo_transcript_script(const char* proto, const uint8_t* script, size_t script_len, uint8_t* out) {
    Transcript tr(proto);
    size_t p = 0;
    while (p < script_len) {
        uint8_t kind = script[p++];
        uint32_t ll, dl;
        memcpy(&ll, script + p, 4);
        p += 4;
        std::string label((const char*)script + p, ll);
        p += ll;
        memcpy(&dl, script + p, 4);
        p += 4;
        if (kind == 0) {
            tr.append_message(label.c_str(), script + p, dl);
            p += dl;
        } else {
            tr.challenge_bytes(label.c_str(), out, dl);
            out += dl;
        }
    }
}

// ---- ark-serialize 0.3 bytes of Proof<Fr, KZG10<Bls12_381>> (proof.rs:37-121; derive order = declaration order) ----
// commitments: compressed G1; openings: kzg10::Proof { w, random_v: None } = 48 B + 1 B; evaluations: wire, perm, lookup
// structs field by field, custom evals as Vec<(String, F)>.
size_t zpo_proof_serialize(const uint64_t* proof_in, uint8_t* out) {
    ensure_init();
    ProofO p = proof_from_bytes(proof_in);
    uint8_t* o = out;
    for (int c = 0; c < 17; c++, o += 48) serialize_g1(p.comm[c], o);
    for (int c = 17; c < 19; c++) {
        serialize_g1(p.comm[c], o);
        o += 48;
        *o++ = 0;  // Option::None
    }
    for (int e = 0; e < 16; e++, o += 32) p.eval[e].to_bytes_le(o);
    static const char* labels[10] = {"q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "q_hl_eval",
                                     "q_hr_eval", "q_h4_eval", "a_next_eval", "b_next_eval", "d_next_eval"};
    uint64_t cnt = 10;
    memcpy(o, &cnt, 8);
    o += 8;
    for (int k = 0; k < 10; k++) {
        uint64_t len = strlen(labels[k]);
        memcpy(o, &len, 8);
        o += 8;
        memcpy(o, labels[k], len);
        o += len;
        p.eval[16 + k].to_bytes_le(o);
        o += 32;
    }
    return (size_t)(o - out);
}

// ---- sigma permutations from a wire map (permutation/mod.rs:101-166): vars = 4 columns (left, right, output, fourth)
// of n_gates variable ids added gate by gate with add_variables_to_map.  code[k * N + i] = (wire << 28) | gate of
// sigma_k(i) (WireData), enc = its field encoding K_wire * omega^gate (compute_permutation_lagrange), N = 2^logn.
void zpo_sigma_from_wires(int logn, size_t n_gates, const uint32_t* vars, size_t n_vars, uint32_t* code, uint64_t* enc) {
    ensure_init();
    Composer cs;
    for (size_t v = 0; v < n_vars; v++) cs.add_input(Fr::zero());
    for (size_t g = 0; g < n_gates; g++) cs.push_gate(vars[g], vars[n_gates + g], vars[2 * n_gates + g], vars[3 * n_gates + g]);
    Domain dom(logn);
    std::vector<Fr> sigma[4];
    compute_sigma_evals(cs, dom, sigma);
    for (int k = 0; k < 4; k++)
        for (size_t i = 0; i < dom.n; i++) {
            memcpy(enc + 4 * (k * dom.n + i), sigma[k][i].v, 32);
            // decode: which (wire, gate) has this encoding
            uint32_t c = 0xffffffffu;
            for (int w = 0; w < 4 && c == 0xffffffffu; w++)
                for (size_t j = 0; j < dom.n; j++)
                    if (K_const(w) * dom.element(j) == sigma[k][i]) {
                        c = ((uint32_t)w << 28) | (uint32_t)j;
                        break;
                    }
            code[k * dom.n + i] = c;
        }
}
// util.rs:154-176 lc(values, challenge) for k scalars
void zpo_lc(size_t k, const uint64_t* values, const uint64_t* challenge, uint64_t* out) {
    ensure_init();
    Fr ch, acc;
    memcpy(ch.v, challenge, 32);
    memcpy(acc.v, values + 4 * (k - 1), 32);
    for (size_t i = k - 1; i-- > 0;) {
        Fr v;
        memcpy(v.v, values + 4 * i, 32);
        acc = acc * ch + v;
    }
    memcpy(out, acc.v, 32);
}
// combine_split for multisets of different cardinality; returns 0 on ElementNotIndexed
int zpo_multiset_combine_split(size_t nt, const uint64_t* t, size_t nf, const uint64_t* f, uint64_t* h1, uint64_t* h2) {
    ensure_init();
    std::vector<Fr> tv(nt), fv(nf), a, b;
    memcpy(tv.data(), t, 32 * nt);
    if (nf) memcpy(fv.data(), f, 32 * nf);
    if (!combine_split(tv, fv, a, b)) return 0;
    memcpy(h1, a.data(), 32 * a.size());
    if (!b.empty()) memcpy(h2, b.data(), 32 * b.size());
    return 1;
}

// ---- combine_split (multiset.rs:131-176) ------------------------------------------------------
int zpo_combine_split(size_t n, const uint64_t* t, const uint64_t* f, uint64_t* h1, uint64_t* h2) {
    ensure_init();
    std::vector<Fr> tv(n), fv(n), a, b;
    memcpy(tv.data(), t, 32 * n);
    memcpy(fv.data(), f, 32 * n);
    if (!combine_split(tv, fv, a, b) || a.size() != n || b.size() != n) return 0;
    memcpy(h1, a.data(), 32 * n);
    memcpy(h2, b.data(), 32 * n);
    return 1;
}

// ---- circuit / prover / verifier context --------------------------------------------------------
// with_pk: 0 = circuit + selector evaluations only (pk built elsewhere), 1 = full CPU preprocessing
// kind: 0 Poseidon-Merkle tree of `height`; 1..3 the gadget circuits of build_custom_circuit (kind 1: `height` = number of
// repetitions of the gadget block, 0 or 1 = once)
void* zpo_ctx_new_kind(int kind, int height, uint64_t witness_seed, uint64_t tau_seed, int n_lookup, int with_pk, int with_srs) {
    ensure_init();
    OracleCtx* c = new OracleCtx();
    c->cs = kind == 0 ? build_merkle_circuit(height, witness_seed, n_lookup) : build_custom_circuit(kind, witness_seed, n_lookup, height);
    size_t bound = std::max(c->cs.n(), c->cs.table.size());
    int logn = log2_ceil(bound);
    size_t N = (size_t)1 << logn;
    c->pk.logn = logn;
    c->pk.n = N;
    for (int k = 0; k < 4; k++) {
        c->w[k].resize(c->cs.n());
        for (size_t i = 0; i < c->cs.n(); i++) c->w[k][i] = c->cs.var_vals[c->cs.w[k][i]];
    }
    if (with_pk) {
        c->pk = preprocess(c->cs);
        c->have_pk = true;
    }
    {
        Domain dom(logn);
        std::vector<Fr> sigma[4];
        compute_sigma_evals(c->cs, dom, sigma);
        for (int s = 0; s < NUM_PK_POLYS; s++) {
            c->sel_evals[s] = s < NUM_SELECTORS ? c->cs.q[s] : sigma[s - NUM_SELECTORS];
            c->sel_evals[s].resize(N, Fr::zero());
        }
        if (!with_pk) {
            for (int col = 0; col < 4; col++) {
                std::vector<Fr>& t = c->pk.table[col];
                for (auto& row : c->cs.table) t.push_back(row[col]);
                if (t.empty()) t.push_back(Fr::zero());
                t.resize(N, t[0]);
            }
        }
    }
    SplitMix64 rng(tau_seed);
    c->tau = rng.next_fr();
    if (with_srs) {
        c->srs = srs_from_tau(c->tau, N);
        c->srs_raw.resize(12 * N);
        for (size_t i = 0; i < N; i++) {
            memcpy(&c->srs_raw[12 * i], c->srs[i].x.v, 48);
            memcpy(&c->srs_raw[12 * i + 6], c->srs[i].y.v, 48);
        }
    }
    // the FFI carries exactly one public input (CircuitC.pi / intended_pi_pos); a circuit without one passes pi = 0,
    // which PublicInputs drops (pi.rs:55-62)
    assert(c->cs.pi.size() <= 1);
    c->pi_pos = 0;
    memset(c->pi_canonical, 0, 32);
    if (c->cs.pi.size() == 1) {
        c->pi_pos = c->cs.pi[0].first;
        c->cs.pi[0].second.to_canonical(c->pi_canonical);
    }
    return c;
}
void* zpo_ctx_new(int height, uint64_t witness_seed, uint64_t tau_seed, int n_lookup, int with_pk, int with_srs) {
    return zpo_ctx_new_kind(0, height, witness_seed, tau_seed, n_lookup, with_pk, with_srs);
}
void zpo_ctx_free(void* h) { delete (OracleCtx*)h; }
uint64_t zpo_ctx_n(void* h) { return ((OracleCtx*)h)->cs.n(); }
int zpo_ctx_logn(void* h) { return ((OracleCtx*)h)->pk.logn; }
uint64_t zpo_ctx_lookup_len(void* h) { return ((OracleCtx*)h)->cs.table.size(); }
uint64_t zpo_ctx_pi_pos(void* h) { return ((OracleCtx*)h)->pi_pos; }
const uint64_t* zpo_ctx_pi(void* h) { return ((OracleCtx*)h)->pi_canonical; }
const uint64_t* zpo_ctx_wire(void* h, int k) { return (const uint64_t*)((OracleCtx*)h)->w[k].data(); }
const uint64_t* zpo_ctx_q_lookup(void* h) { return (const uint64_t*)((OracleCtx*)h)->cs.q[Q_LOOKUP].data(); }
const uint64_t* zpo_ctx_selector_evals(void* h, int s) { return (const uint64_t*)((OracleCtx*)h)->sel_evals[s].data(); }
const uint64_t* zpo_ctx_pk_coeffs(void* h, int s) { return (const uint64_t*)((OracleCtx*)h)->pk.coeffs[s].data(); }
const uint64_t* zpo_ctx_pk_evals(void* h, int s) { return (const uint64_t*)((OracleCtx*)h)->pk.evals[s].data(); }
const uint64_t* zpo_ctx_table(void* h, int c) { return (const uint64_t*)((OracleCtx*)h)->pk.table[c].data(); }
const uint64_t* zpo_ctx_linear_evaluations(void* h) { return (const uint64_t*)((OracleCtx*)h)->pk.linear_evaluations.data(); }
const uint64_t* zpo_ctx_v_h_coset_8n(void* h) { return (const uint64_t*)((OracleCtx*)h)->pk.v_h_coset_8n.data(); }
const uint64_t* zpo_ctx_srs(void* h) { return ((OracleCtx*)h)->srs_raw.data(); }
const uint64_t* zpo_ctx_tau(void* h) { return ((OracleCtx*)h)->tau.v; }
// the composer's wire map in insertion order (Permutation::variable_map flattened): m entries (variable, (gate << 2) | wire)
uint64_t zpo_ctx_wiring_len(void* h) { return ((OracleCtx*)h)->cs.perm_log.size(); }
uint64_t zpo_ctx_num_vars(void* h) { return ((OracleCtx*)h)->cs.var_vals.size(); }
void zpo_ctx_wiring(void* h, uint32_t* vars, uint32_t* cells) {
    const auto& log = ((OracleCtx*)h)->cs.perm_log;
    for (size_t i = 0; i < log.size(); i++) {
        vars[i] = log[i].first;
        cells[i] = log[i].second;
    }
}

// inputs of build_merkle_circuit for an external witness generator: the 8 blinding values, the 2^(height-1) leaves (both
// drawn from the witness seed in the generator's order) and the hash parameters (9 MDS row-major, 3 pre-round keys, 63 x 3
// round constants)
void zpo_merkle_inputs(int height, uint64_t witness_seed, uint64_t* blinding8, uint64_t* leaves, uint64_t* params201) {
    ensure_init();
    SplitMix64 rng(witness_seed);
    for (int i = 0; i < 8; i++) memcpy(blinding8 + 4 * i, rng.next_fr().v, 32);
    size_t nleaves = (size_t)1 << (height - 1);
    for (size_t i = 0; i < nleaves; i++) memcpy(leaves + 4 * i, rng.next_fr().v, 32);
    HashParams hp(0x504f534549444f4eULL);
    uint64_t* o = params201;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++, o += 4) memcpy(o, hp.mds[i][j].v, 32);
    for (int j = 0; j < 3; j++, o += 4) memcpy(o, hp.ark0[j].v, 32);
    for (int r = 0; r < 63; r++)
        for (int j = 0; j < 3; j++, o += 4) memcpy(o, hp.rc[r][j].v, 32);
}

// every gate equation of the synthetic circuit holds on the witness (sanity of the generator itself): arithmetic + PI
// + the four custom widgets (separation challenges drawn at random) on every row, "next" = the following row, and
// every lookup row is in the table
int zpo_ctx_check_satisfied(void* h) {
    OracleCtx* c = (OracleCtx*)h;
    Composer& cs = c->cs;
    const size_t n = cs.n();
    std::vector<Fr> pi(n, Fr::zero());
    for (auto& e : cs.pi) pi[e.first] = e.second;
    SplitMix64 rng(0x5e9a);
    Fr rs = rng.next_fr(), ls = rng.next_fr(), fs = rng.next_fr(), vs = rng.next_fr();
    auto wv = [&](int k, size_t i) { return i < n ? c->w[k][i] : Fr::zero(); };
    for (size_t i = 0; i < n; i++) {
        GateVals g;
        g.a = wv(0, i); g.b = wv(1, i); g.c = wv(2, i); g.d = wv(3, i);
        g.a_next = wv(0, i + 1); g.b_next = wv(1, i + 1); g.d_next = wv(3, i + 1);
        g.q_l = cs.q[Q_L][i]; g.q_r = cs.q[Q_R][i]; g.q_c = cs.q[Q_C][i];
        Fr v = (g.a * g.b * cs.q[Q_M][i] + g.a * cs.q[Q_L][i] + g.b * cs.q[Q_R][i] + g.c * cs.q[Q_O][i] + g.d * cs.q[Q_4][i] +
                g.a.pow_u64(5) * cs.q[Q_HL][i] + g.b.pow_u64(5) * cs.q[Q_HR][i] + g.d.pow_u64(5) * cs.q[Q_H4][i] + cs.q[Q_C][i]) *
                   cs.q[Q_ARITH][i] +
               pi[i];
        v = v + cs.q[Q_RANGE][i] * range_constraints(rs, g) + cs.q[Q_LOGIC][i] * logic_constraints(ls, g) +
            cs.q[Q_FIXED][i] * fbsm_constraints(fs, g) + cs.q[Q_VAR][i] * curve_add_constraints(vs, g);
        if (!v.is_zero()) return 0;
        if (!cs.q[Q_LOOKUP][i].is_zero()) {
            bool found = false;
            for (auto& row : cs.table)
                if (row[0] == g.a && row[1] == g.b && row[2] == g.c && row[3] == g.d) found = true;
            if (!found) return 0;
        }
    }
    // copy constraints: every cell of a variable's cycle holds the variable's value
    for (auto& e : cs.perm_log)
        if (!(c->w[e.second & 3][e.second >> 2] == cs.var_vals[e.first])) return 0;
    return 1;
}
int zpo_te_generator_on_curve() {
    ensure_init();
    return te_on_curve(te_generator()) && (-jubjub_d() == Fr::from_u64(10240) * Fr::from_u64(10241).inverse()) ? 1 : 0;
}

// Runs the CPU restatement of the prover; proof_out = 2656-byte ProofC image; returns seconds.
double zpo_ctx_prove(void* h, uint64_t* proof_out, uint64_t* challenges_out) {
    OracleCtx* c = (OracleCtx*)h;
    assert(c->have_pk && !c->srs.empty());
    ProverInput in;
    in.pk = &c->pk;
    in.srs = &c->srs;
    for (int k = 0; k < 4; k++) in.w[k] = c->w[k];
    in.q_lookup = c->cs.q[Q_LOOKUP];
    in.pi = c->cs.pi;
    in.label = c->label;
    Challenges ch;
    auto t0 = std::chrono::steady_clock::now();
    ProofO p = prove(in, &ch);
    auto t1 = std::chrono::steady_clock::now();
    proof_to_bytes(p, proof_out);
    if (challenges_out) memcpy(challenges_out, &ch, sizeof(ch));
    return std::chrono::duration<double>(t1 - t0).count();
}
// Verifier restatement on a ProofC image. Returns 1 accept / 0 reject; bit0 aw, bit1 saw in *detail.
int zpo_ctx_verify(void* h, const uint64_t* proof_in, int* detail) {
    OracleCtx* c = (OracleCtx*)h;
    if (!c->have_vk) {
        assert(c->have_pk && !c->srs.empty());
        c->vk = make_verifier_key(c->pk, c->srs);
        c->have_vk = true;
    }
    ProofO p = proof_from_bytes(proof_in);
    VerifyTrace t;
    bool ok = verify(c->vk, p, c->cs.pi, c->label, c->tau, &t);
    if (detail) *detail = (t.aw_ok ? 1 : 0) | (t.saw_ok ? 2 : 0);
    return ok ? 1 : 0;
}
// Verifier key built from externally supplied commitments (e.g. produced by the GPU path at sizes
// where the CPU MSM is too slow): 19 pk commitments + 4 table commitments, 12 u64 each.
void zpo_ctx_set_vk(void* h, const uint64_t* comms23) {
    OracleCtx* c = (OracleCtx*)h;
    c->vk.logn = c->pk.logn;
    c->vk.n = c->pk.n;
    for (int s = 0; s < NUM_PK_POLYS + 4; s++) {
        G1Affine a;
        memcpy(a.x.v, comms23 + 12 * s, 48);
        memcpy(a.y.v, comms23 + 12 * s + 6, 48);
        a.inf = a.x.is_zero() && a.y == Fq::one();
        if (s < NUM_PK_POLYS)
            c->vk.pk_comm[s] = a;
        else
            c->vk.table_comm[s - NUM_PK_POLYS] = a;
    }
    c->have_vk = true;
}
// commit(p) with the known trapdoor: [p(tau)] G  (one scalar multiplication)
void zpo_ctx_commit_with_tau(void* h, size_t n, const uint64_t* coeffs, uint64_t* out) {
    OracleCtx* c = (OracleCtx*)h;
    std::vector<Fr> p(n);
    memcpy(p.data(), coeffs, 32 * n);
    Fr v = poly_eval(p, c->tau);
    G1Affine r = G1::from_affine(g1_generator()).mul(v).to_affine();
    memcpy(out, r.x.v, 48);
    memcpy(out + 6, r.y.v, 48);
}

// ---- timing helpers for the CPU baseline (bench.py cpu_baseline / --impl reference) ---------------
double zpo_time_ntt(int kind, int logn, int iters, uint64_t seed) {
    Domain dom(logn);
    std::vector<Fr> a(dom.n);
    zpo_random_fr(seed, dom.n, (uint64_t*)a.data());
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; i++) a = kind == 0 ? dom.fft(a) : kind == 1 ? dom.ifft(a) : kind == 2 ? dom.coset_fft(a) : dom.coset_ifft(a);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count() / iters;
}
double zpo_time_msm(size_t n, const uint64_t* points, const uint64_t* scalars, uint64_t* out) {
    auto t0 = std::chrono::steady_clock::now();
    zpo_msm(n, points, scalars, out);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
