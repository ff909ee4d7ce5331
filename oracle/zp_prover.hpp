// ORACLE — TEST INFRASTRUCTURE ONLY (see zp_field.hpp header).
//
// CPU restatement of the ZK-Garage prover `Prover::prove_with_preprocessed`
// ("Prize 1B/plonk-core/src/proof_system/prover.rs":171-660), function by function:
//   rounds / transcript schedule ......... prover.rs:187-636
//   MultiSet::compress / lc .............. lookup/multiset.rs:207-213, util.rs:154-176
//   MultiSet::combine_split .............. lookup/multiset.rs:131-176
//   compute_permutation_poly ............. permutation/mod.rs:652-752
//   compute_lookup_permutation_poly ...... permutation/mod.rs:754-838
//   quotient_poly::compute ............... proof_system/quotient_poly.rs:34-358
//   gate widgets ......................... proof_system/widget/{arithmetic.rs:61-79, range.rs:43-74,
//                                          logic.rs:60-133, ecc/fixed_base_scalar_mul.rs:82-156,
//                                          ecc/curve_addition.rs:52-97, lookup.rs:42-215}
//   permutation widget ................... proof_system/permutation.rs:62-296
//   linearisation_poly::compute .......... proof_system/linearisation_poly.rs:164-432
//   KZG commit / open (SonicKZG10, no hiding)  PNP twin lib/PLONK/src/KZG/kzg10.cu:31-146
// Parity: the Rust crate cannot be built here (no cargo; arkworks/merlin not vendored), so this
// restatement is pinned by (1) the reference's own constants, (2) blst + the reference's strobe.cpp
// compiled into oracle/_ref, (3) Merlin's published known-answer vector, (4) the reference's own
// unit-test vectors for combine_split, the sigma permutations and lc (tests/test_reference_vectors.py),
// (5) the reference's NATIVE prover run on the GPU box at HEIGHT=4 (same ProofC bytes), (6) two
// verifiers accepting its proofs: the trapdoor restatement (zp_verifier.hpp) and the product's pairing
// verifier.  No reference fixture holds proof bytes, so whole-proof parity against the
// RUST prover itself remains "unpinned".
#pragma once
#include "zp_circuit.hpp"
#include "zp_transcript.hpp"
#include <unordered_map>
#include <cstdio>
#include <cstdlib>
#include <chrono>

namespace zpo {

// Order of ProofC commitments (lib.rs:120-142) and ProofEvaluationsC scalars (lib.rs:53-118).
enum ProofComm { C_A = 0, C_B, C_C, C_D, C_Z, C_F, C_H1, C_H2, C_Z2, C_T1, C_T2, C_T3, C_T4, C_T5, C_T6, C_T7, C_T8, C_AW, C_SAW, NUM_COMM };
enum ProofEval {
    E_A = 0, E_B, E_C, E_D,
    E_LSIG, E_RSIG, E_OSIG, E_PERM,
    E_QLOOKUP, E_Z2NEXT, E_H1, E_H1NEXT, E_H2, E_F, E_TABLE, E_TABLENEXT,
    E_QARITH, E_QC, E_QL, E_QR, E_QHL, E_QHR, E_QH4, E_ANEXT, E_BNEXT, E_DNEXT,
    NUM_EVAL
};

struct ProofO {
    G1Affine comm[NUM_COMM];
    Fr eval[NUM_EVAL];
};

struct Challenges {
    Fr zeta, beta, gamma, delta, epsilon, alpha, range_sep, logic_sep, fixed_sep, var_sep, lookup_sep, z, aw, saw;
};

// JubJub (ed-on-bls12-381) parameters for the ECC gates: a = -1, d = -(10240/10241)
// (in-tree Montgomery literals: "…/lib/PLONK/src/bls12_381/edwards.cu":5-31).
static inline Fr jubjub_a() { return -Fr::one(); }
static inline const Fr& jubjub_d() {
    static const Fr d = -(Fr::from_u64(10240) * Fr::from_u64(10241).inverse());
    return d;
}
// Montgomery forms of the small integers the widgets use, converted once
static inline const Fr& small_fr(unsigned k) {
    static const std::vector<Fr> tab = [] {
        ensure_init();
        std::vector<Fr> t(128);
        for (unsigned i = 0; i < 128; i++) t[i] = Fr::from_u64(i);
        return t;
    }();
    return tab[k];
}

static inline Fr lc4(const Fr& a, const Fr& b, const Fr& c, const Fr& d, const Fr& ch) {
    return ((d * ch + c) * ch + b) * ch + a;  // util.rs:170-175
}
static inline Fr delta4(const Fr& f) {  // f(f-1)(f-2)(f-3)
    return f * (f - small_fr(1)) * (f - small_fr(2)) * (f - small_fr(3));
}

struct GateVals {
    Fr a, b, c, d, a_next, b_next, d_next, q_l, q_r, q_c;
};
static inline Fr range_constraints(const Fr& sep, const GateVals& g) {  // range.rs:49-63
    const Fr& four = small_fr(4);
    Fr kappa = sep.square(), kappa_sq = kappa.square(), kappa_cu = kappa_sq * kappa;
    Fr b1 = delta4(g.c - four * g.d);
    Fr b2 = delta4(g.b - four * g.c) * kappa;
    Fr b3 = delta4(g.a - four * g.b) * kappa_sq;
    Fr b4 = delta4(g.d_next - four * g.a) * kappa_cu;
    return (b1 + b2 + b3 + b4) * sep;
}
static inline Fr delta_xor_and(const Fr& a, const Fr& b, const Fr& w, const Fr& c, const Fr& q_c) {  // logic.rs:113-133
    const Fr &nine = small_fr(9), &two = small_fr(2), &three = small_fr(3), &four = small_fr(4);
    const Fr &eighteen = small_fr(18), &eighty_one = small_fr(81), &eighty_three = small_fr(83);
    Fr F = w * (w * (four * w - eighteen * (a + b) + eighty_one) + eighteen * (a.square() + b.square()) -
                eighty_one * (a + b) + eighty_three);
    Fr E = three * (a + b + c) - (two * F);
    Fr B = q_c * ((nine * c) - three * (a + b));
    return B + E;
}
static inline Fr logic_constraints(const Fr& sep, const GateVals& g) {  // logic.rs:66-91
    const Fr& four = small_fr(4);
    Fr kappa = sep.square(), kappa_sq = kappa.square(), kappa_cu = kappa_sq * kappa, kappa_qu = kappa_cu * kappa;
    Fr a = g.a_next - four * g.a;
    Fr c0 = delta4(a);
    Fr b = g.b_next - four * g.b;
    Fr c1 = delta4(b) * kappa;
    Fr d = g.d_next - four * g.d;
    Fr c2 = delta4(d) * kappa_sq;
    Fr w = g.c;
    Fr c3 = (w - a * b) * kappa_cu;
    Fr c4 = delta_xor_and(a, b, w, d, g.q_c) * kappa_qu;
    return (c0 + c1 + c2 + c3 + c4) * sep;
}
static inline Fr fbsm_constraints(const Fr& sep, const GateVals& g) {  // fixed_base_scalar_mul.rs:88-139
    Fr kappa = sep.square(), kappa_sq = kappa.square(), kappa_cu = kappa_sq * kappa;
    Fr x_beta = g.q_l, y_beta = g.q_r;
    Fr acc_x = g.a, acc_x_next = g.a_next, acc_y = g.b, acc_y_next = g.b_next;
    Fr xy_alpha = g.c;
    Fr bit = g.d_next - g.d - g.d;
    Fr bit_consistency = bit * (bit - Fr::one()) * (bit + Fr::one());
    Fr y_alpha = bit.square() * (y_beta - Fr::one()) + Fr::one();
    Fr x_alpha = x_beta * bit;
    Fr xy_consistency = ((bit * g.q_c) - xy_alpha) * kappa;
    Fr x_3 = acc_x_next;
    Fr lhs = x_3 + (x_3 * xy_alpha * acc_x * acc_y * jubjub_d());
    Fr rhs = (x_alpha * acc_y) + (y_alpha * acc_x);
    Fr x_acc = (lhs - rhs) * kappa_sq;
    Fr y_3 = acc_y_next;
    lhs = y_3 - (y_3 * xy_alpha * acc_x * acc_y * jubjub_d());
    rhs = y_alpha * acc_y - jubjub_a() * x_alpha * acc_x;
    Fr y_acc = (lhs - rhs) * kappa_cu;
    return (bit_consistency + x_acc + y_acc + xy_consistency) * sep;
}
static inline Fr curve_add_constraints(const Fr& sep, const GateVals& g) {  // curve_addition.rs:59-95
    Fr x_1 = g.a, x_3 = g.a_next, y_1 = g.b, y_3 = g.b_next, x_2 = g.c, y_2 = g.d, x1_y2 = g.d_next;
    Fr kappa = sep.square();
    Fr xy_consistency = x_1 * y_2 - x1_y2;
    Fr y1_x2 = y_1 * x_2, y1_y2 = y_1 * y_2, x1_x2 = x_1 * x_2;
    Fr x3_lhs = x1_y2 + y1_x2;
    Fr x3_rhs = x_3 + (x_3 * jubjub_d() * x1_y2 * y1_x2);
    Fr x3_consistency = (x3_lhs - x3_rhs) * kappa;
    Fr y3_lhs = y1_y2 - jubjub_a() * x1_x2;
    Fr y3_rhs = y_3 - y_3 * jubjub_d() * x1_y2 * y1_x2;
    Fr y3_consistency = (y3_lhs - y3_rhs) * kappa.square();
    return (xy_consistency + x3_consistency + y3_consistency) * sep;
}

// lookup/multiset.rs:131-176 (IndexMap = insertion-ordered buckets)
static inline bool combine_split(const std::vector<Fr>& t, const std::vector<Fr>& f, std::vector<Fr>& h1, std::vector<Fr>& h2) {
    struct KeyHash {
        size_t operator()(const std::array<uint64_t, 4>& k) const { return k[0] ^ (k[1] * 0x9e3779b97f4a7c15ULL) ^ (k[2] << 1) ^ (k[3] >> 1); }
    };
    std::unordered_map<std::array<uint64_t, 4>, size_t, KeyHash> idx;
    std::vector<Fr> keys;
    std::vector<size_t> counts;
    auto key_of = [](const Fr& x) { return std::array<uint64_t, 4>{x.v[0], x.v[1], x.v[2], x.v[3]}; };
    for (auto& e : t) {
        auto it = idx.find(key_of(e));
        if (it == idx.end()) {
            idx[key_of(e)] = keys.size();
            keys.push_back(e);
            counts.push_back(1);
        } else
            counts[it->second]++;
    }
    for (auto& e : f) {
        auto it = idx.find(key_of(e));
        if (it == idx.end()) return false;  // Error::ElementNotIndexed
        counts[it->second]++;
    }
    h1.clear();
    h2.clear();
    int parity = 0;
    for (size_t k = 0; k < keys.size(); k++) {
        size_t half = counts[k] / 2;
        h1.insert(h1.end(), half, keys[k]);
        h2.insert(h2.end(), half, keys[k]);
        if (counts[k] % 2 == 1) {
            if (parity == 1) {
                h2.push_back(keys[k]);
                parity = 0;
            } else {
                h1.push_back(keys[k]);
                parity = 1;
            }
        }
    }
    return true;
}

struct ProverInput {
    const ProverKeyO* pk;
    const std::vector<G1Affine>* srs;  // powers_of_g, at least N points
    std::vector<Fr> w[4];              // unpadded witness columns (cs.n entries)
    std::vector<Fr> q_lookup;          // unpadded
    std::vector<std::pair<uint64_t, Fr>> pi;
    std::string label;
};

static inline G1Affine kzg_commit(const std::vector<G1Affine>& srs, const std::vector<Fr>& coeffs) {
    assert(coeffs.size() <= srs.size());
    return g1_msm(srs.data(), coeffs.data(), coeffs.size()).to_affine();
}

// open: p(X) = sum_j ch^j p_j(X); W = commit(floor(p / (X - point)))
static inline G1Affine kzg_open(const std::vector<G1Affine>& srs, const std::vector<const std::vector<Fr>*>& polys,
                                const Fr& point, const Fr& ch) {
    size_t n = 0;
    for (auto p : polys) n = std::max(n, p->size());
    std::vector<Fr> comb(n, Fr::zero());
    Fr cj = Fr::one();
    for (auto p : polys) {
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (long i = 0; i < (long)p->size(); i++) comb[i] = comb[i] + (*p)[i] * cj;
        cj = cj * ch;
    }
    std::vector<Fr> w = poly_div_linear(comb, point);
    return kzg_commit(srs, w);
}

struct PhaseClock {  // ZPO_DEBUG=1: wall time per protocol phase on stderr
    bool on = getenv("ZPO_DEBUG") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[zpo] %-28s %8.3f s\n", what, std::chrono::duration<double>(n - t).count());
        t = n;
    }
};

static inline ProofO prove(const ProverInput& in, Challenges* ch_out = nullptr) {
    ensure_init();
    PhaseClock clk;
    const ProverKeyO& pk = *in.pk;
    const std::vector<G1Affine>& srs = *in.srs;
    const size_t n = pk.n, n8 = 8 * n;
    Domain dom(pk.logn), dom8(pk.logn + 3);
    ProofO proof;
    Challenges ch;
    Transcript tr(in.label);
    tr.append_pi("pi", in.pi);

    // 1. witness polynomials
    std::vector<Fr> w_ev[4], w_poly[4];
    for (int k = 0; k < 4; k++) {
        w_ev[k] = in.w[k];
        w_ev[k].resize(n, Fr::zero());
        w_poly[k] = dom.ifft(w_ev[k]);
        proof.comm[C_A + k] = kzg_commit(srs, w_poly[k]);
    }
    tr.append_g1("w_l", proof.comm[C_A]);
    tr.append_g1("w_r", proof.comm[C_B]);
    tr.append_g1("w_o", proof.comm[C_C]);
    tr.append_g1("w_4", proof.comm[C_D]);

    clk.lap("1 wires: 4 ifft + 4 commit");
    // 2. lookup polynomials
    ch.zeta = tr.challenge_scalar("zeta");
    tr.append_fr("zeta", ch.zeta);
    std::vector<Fr> t_ev(n), f_ev(n);
    for (size_t i = 0; i < n; i++) t_ev[i] = lc4(pk.table[0][i], pk.table[1][i], pk.table[2][i], pk.table[3][i], ch.zeta);
    std::vector<Fr> table_poly = dom.ifft(t_ev);
    for (size_t i = 0; i < n; i++) {
        bool on = i < in.q_lookup.size() && !in.q_lookup[i].is_zero();
        if (on)
            f_ev[i] = lc4(w_ev[0][i], w_ev[1][i], w_ev[2][i], w_ev[3][i], ch.zeta);
        else
            f_ev[i] = lc4(t_ev[0], Fr::zero(), Fr::zero(), Fr::zero(), ch.zeta);
    }
    std::vector<Fr> f_poly = dom.ifft(f_ev);
    proof.comm[C_F] = kzg_commit(srs, f_poly);
    tr.append_g1("f", proof.comm[C_F]);
    std::vector<Fr> h1_ev, h2_ev;
    bool ok = combine_split(t_ev, f_ev, h1_ev, h2_ev);
    assert(ok && h1_ev.size() == n && h2_ev.size() == n);
    (void)ok;
    std::vector<Fr> h1_poly = dom.ifft(h1_ev), h2_poly = dom.ifft(h2_ev);
    proof.comm[C_H1] = kzg_commit(srs, h1_poly);
    proof.comm[C_H2] = kzg_commit(srs, h2_poly);
    tr.append_g1("h1", proof.comm[C_H1]);
    tr.append_g1("h2", proof.comm[C_H2]);

    clk.lap("2 lookup polys");
    // 3. permutation polynomials
    ch.beta = tr.challenge_scalar("beta");
    tr.append_fr("beta", ch.beta);
    ch.gamma = tr.challenge_scalar("gamma");
    tr.append_fr("gamma", ch.gamma);
    ch.delta = tr.challenge_scalar("delta");
    tr.append_fr("delta", ch.delta);
    ch.epsilon = tr.challenge_scalar("epsilon");
    tr.append_fr("epsilon", ch.epsilon);

    std::vector<Fr> z_poly;
    {
        std::vector<Fr> sig[4];
        for (int k = 0; k < 4; k++) sig[k] = dom.fft(pk.coeffs[SIG_L + k]);
        std::vector<Fr> num(n), den(n);
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (long i = 0; i < (long)n; i++) {
            Fr root = dom.element(i);
            Fr nu = Fr::one(), de = Fr::one();
            for (int k = 0; k < 4; k++) {
                nu = nu * (w_ev[k][i] + ch.beta * K_const(k) * root + ch.gamma);
                de = de * (w_ev[k][i] + ch.beta * sig[k][i] + ch.gamma);
            }
            num[i] = nu;
            den[i] = de;
        }
        batch_inverse(den);
        std::vector<Fr> z(n);
        Fr state = Fr::one();
        for (size_t i = 0; i < n; i++) {
            z[i] = state;
            state = state * (num[i] * den[i]);
        }
        if (getenv("ZPO_DEBUG")) fprintf(stderr, "[zpo] z closes: %d\n", (int)(state == Fr::one()));
        z_poly = dom.ifft(z);
    }
    proof.comm[C_Z] = kzg_commit(srs, z_poly);
    tr.append_g1("z", proof.comm[C_Z]);

    std::vector<Fr> z2_poly;
    {
        Fr opd = Fr::one() + ch.delta, eopd = ch.epsilon * opd;
        std::vector<Fr> num(n), den(n);
        for (size_t i = 0; i < n; i++) {
            size_t nx = (i + 1) % n;
            num[i] = opd * (ch.epsilon + f_ev[i]) * (eopd + t_ev[i] + ch.delta * t_ev[nx]);
            den[i] = (eopd + h1_ev[i] + h2_ev[i] * ch.delta) * (eopd + h2_ev[i] + h1_ev[nx] * ch.delta);
        }
        batch_inverse(den);
        std::vector<Fr> p(n);
        Fr state = Fr::one();
        for (size_t i = 0; i < n; i++) {
            p[i] = state;
            state = state * (num[i] * den[i]);
        }
        z2_poly = dom.ifft(p);
    }
    proof.comm[C_Z2] = kzg_commit(srs, z2_poly);  // NOT appended to the transcript (prover.rs:395-397)

    clk.lap("3 z, z2");
    std::vector<Fr> pi_ev(n, Fr::zero());
    for (auto& e : in.pi) pi_ev[e.first] = e.second;
    std::vector<Fr> pi_poly = dom.ifft(pi_ev);

    // 4. quotient
    ch.alpha = tr.challenge_scalar("alpha");
    tr.append_fr("alpha", ch.alpha);
    ch.range_sep = tr.challenge_scalar("range separation challenge");
    tr.append_fr("range seperation challenge", ch.range_sep);
    ch.logic_sep = tr.challenge_scalar("logic separation challenge");
    tr.append_fr("logic seperation challenge", ch.logic_sep);
    ch.fixed_sep = tr.challenge_scalar("fixed base separation challenge");
    tr.append_fr("fixed base separation challenge", ch.fixed_sep);
    ch.var_sep = tr.challenge_scalar("variable base separation challenge");
    tr.append_fr("variable base separation challenge", ch.var_sep);
    ch.lookup_sep = tr.challenge_scalar("lookup separation challenge");
    tr.append_fr("lookup separation challenge", ch.lookup_sep);

    std::vector<Fr> t_poly;
    {
        std::vector<Fr> l1_ev(n, Fr::zero());
        l1_ev[0] = Fr::one();
        std::vector<Fr> l1_poly = dom.ifft(l1_ev);
        std::vector<Fr> l1_8 = dom8.coset_fft(l1_poly);
        std::vector<Fr> l1a_ev(n, Fr::zero());
        l1a_ev[0] = ch.alpha.square();
        std::vector<Fr> l1a_8 = dom8.coset_fft(dom.ifft(l1a_ev));
        std::vector<Fr> z8 = dom8.coset_fft(z_poly), z28 = dom8.coset_fft(z2_poly);
        std::vector<Fr> w8[4];
        for (int k = 0; k < 4; k++) w8[k] = dom8.coset_fft(w_poly[k]);
        std::vector<Fr> f8 = dom8.coset_fft(f_poly), tb8 = dom8.coset_fft(table_poly);
        std::vector<Fr> h18 = dom8.coset_fft(h1_poly), h28 = dom8.coset_fft(h2_poly);
        std::vector<Fr> pi8 = dom8.coset_fft(pi_poly);
        clk.lap("4a coset ffts");
        std::vector<Fr> quot(n8);
        Fr lsep_sq = ch.lookup_sep.square(), lsep_cu = lsep_sq * ch.lookup_sep;
        Fr opd = ch.delta + Fr::one(), eopd = ch.epsilon * opd;
        // v_h_coset_8n is 8-periodic (g^N * w8^i - 1); the reference inverts every entry
        // (quotient_poly.rs:198-199), the value is the same.
        Fr vh_inv[8];
        for (int k = 0; k < 8; k++) vh_inv[k] = pk.v_h_coset_8n[k].inverse();
#pragma omp parallel for schedule(static) if (n8 >= 4096)
        for (long ii = 0; ii < (long)n8; ii++) {
            size_t i = ii, nx = (i + 8) % n8;
            GateVals g;
            g.a = w8[0][i];
            g.b = w8[1][i];
            g.c = w8[2][i];
            g.d = w8[3][i];
            g.a_next = w8[0][nx];
            g.b_next = w8[1][nx];
            g.d_next = w8[3][nx];
            g.q_l = pk.evals[Q_L][i];
            g.q_r = pk.evals[Q_R][i];
            g.q_c = pk.evals[Q_C][i];
            // arithmetic.rs:61-79
            Fr arith = ((g.a * g.b * pk.evals[Q_M][i]) + (g.a * pk.evals[Q_L][i]) + (g.b * pk.evals[Q_R][i]) +
                        (g.c * pk.evals[Q_O][i]) + (g.d * pk.evals[Q_4][i]) + (g.a.pow_u64(5) * pk.evals[Q_HL][i]) +
                        (g.b.pow_u64(5) * pk.evals[Q_HR][i]) + (g.d.pow_u64(5) * pk.evals[Q_H4][i]) + pk.evals[Q_C][i]) *
                       pk.evals[Q_ARITH][i];
            Fr gate = (arith + pi8[i]) + pk.evals[Q_RANGE][i] * range_constraints(ch.range_sep, g) +
                      pk.evals[Q_LOGIC][i] * logic_constraints(ch.logic_sep, g) +
                      pk.evals[Q_FIXED][i] * fbsm_constraints(ch.fixed_sep, g) +
                      pk.evals[Q_VAR][i] * curve_add_constraints(ch.var_sep, g);
            // permutation.rs:62-153
            Fr x = pk.linear_evaluations[i];
            Fr pa = (g.a + (ch.beta * x) + ch.gamma) * (g.b + (ch.beta * K_const(1) * x) + ch.gamma) *
                    (g.c + (ch.beta * K_const(2) * x) + ch.gamma) * (g.d + (ch.beta * K_const(3) * x) + ch.gamma) * z8[i] *
                    ch.alpha;
            Fr pb = -((g.a + (ch.beta * pk.evals[SIG_L][i]) + ch.gamma) * (g.b + (ch.beta * pk.evals[SIG_R][i]) + ch.gamma) *
                      (g.c + (ch.beta * pk.evals[SIG_O][i]) + ch.gamma) * (g.d + (ch.beta * pk.evals[SIG_4][i]) + ch.gamma) *
                      z8[nx] * ch.alpha);
            Fr pc = (z8[i] - Fr::one()) * l1a_8[i];
            Fr perm = pa + pb + pc;
            // lookup.rs:98-152
            Fr la = pk.evals[Q_LOOKUP][i] * (lc4(g.a, g.b, g.c, g.d, ch.zeta) - f8[i]) * ch.lookup_sep;
            Fr lb = z28[i] * opd * (ch.epsilon + f8[i]) * (eopd + tb8[i] + ch.delta * tb8[nx]) * lsep_sq;
            Fr lcx = -z28[nx] * (eopd + h18[i] + ch.delta * h28[i]) * (eopd + h28[i] + ch.delta * h18[nx]) * lsep_sq;
            Fr ld = (z28[i] - Fr::one()) * l1_8[i] * lsep_cu;
            Fr lookup = la + lb + lcx + ld;
            quot[i] = (gate + perm + lookup) * vh_inv[i & 7];
        }
        clk.lap("4b quotient pass");
        t_poly = dom8.coset_ifft(quot);
        clk.lap("4c coset ifft");
        if (getenv("ZPO_DEBUG")) {
            size_t deg = 0;
            for (size_t i = 0; i < n8; i++)
                if (!t_poly[i].is_zero()) deg = i;
            fprintf(stderr, "[zpo] deg t = %zu (n = %zu, 8n = %zu)\n", deg, n, n8);
        }
    }
    std::vector<Fr> t_i[8];
    for (int k = 0; k < 8; k++) {
        t_i[k].assign(t_poly.begin() + k * n, t_poly.begin() + (k + 1) * n);
        proof.comm[C_T1 + k] = kzg_commit(srs, t_i[k]);
    }
    static const char* tl[8] = {"t_1", "t_2", "t_3", "t_4", "t_5", "t_6", "t_7", "t_8"};
    for (int k = 0; k < 8; k++) tr.append_g1(tl[k], proof.comm[C_T1 + k]);

    clk.lap("4d t commits");
    // 5. linearisation
    ch.z = tr.challenge_scalar("z");
    tr.append_fr("z", ch.z);
    Fr zs = ch.z * dom.omega;
    Fr* e = proof.eval;
    {
        struct EvalJob { int slot; const std::vector<Fr>* poly; bool shifted; };
        const EvalJob jobs[NUM_EVAL] = {
            {E_A, &w_poly[0], false}, {E_B, &w_poly[1], false}, {E_C, &w_poly[2], false}, {E_D, &w_poly[3], false},
            {E_LSIG, &pk.coeffs[SIG_L], false}, {E_RSIG, &pk.coeffs[SIG_R], false}, {E_OSIG, &pk.coeffs[SIG_O], false},
            {E_PERM, &z_poly, true}, {E_QARITH, &pk.coeffs[Q_ARITH], false}, {E_QLOOKUP, &pk.coeffs[Q_LOOKUP], false},
            {E_QC, &pk.coeffs[Q_C], false}, {E_QL, &pk.coeffs[Q_L], false}, {E_QR, &pk.coeffs[Q_R], false},
            {E_ANEXT, &w_poly[0], true}, {E_BNEXT, &w_poly[1], true}, {E_DNEXT, &w_poly[3], true},
            {E_QHL, &pk.coeffs[Q_HL], false}, {E_QHR, &pk.coeffs[Q_HR], false}, {E_QH4, &pk.coeffs[Q_H4], false},
            {E_Z2NEXT, &z2_poly, true}, {E_H1, &h1_poly, false}, {E_H1NEXT, &h1_poly, true}, {E_H2, &h2_poly, false},
            {E_F, &f_poly, false}, {E_TABLE, &table_poly, false}, {E_TABLENEXT, &table_poly, true}};
#pragma omp parallel for schedule(dynamic, 1)
        for (int j = 0; j < NUM_EVAL; j++) e[jobs[j].slot] = poly_eval(*jobs[j].poly, jobs[j].shifted ? zs : ch.z);
    }
    Fr vanishing = dom.evaluate_vanishing(ch.z);
    Fr z_to_n = vanishing + Fr::one();
    Fr l1_eval = vanishing * (Fr::from_u64(n) * (ch.z - Fr::one())).inverse();  // proof.rs:647-658

    std::vector<Fr> lin(n, Fr::zero());
    {
        GateVals g;
        g.a = e[E_A];
        g.b = e[E_B];
        g.c = e[E_C];
        g.d = e[E_D];
        g.a_next = e[E_ANEXT];
        g.b_next = e[E_BNEXT];
        g.d_next = e[E_DNEXT];
        g.q_l = e[E_QL];
        g.q_r = e[E_QR];
        g.q_c = e[E_QC];
        // scalar multipliers per polynomial; r(X) = sum_k s_k * poly_k(X)
        std::vector<std::pair<const std::vector<Fr>*, Fr>> terms;
        Fr qa = e[E_QARITH];
        // arithmetic.rs:83-101 : (q_m ab + q_l a + q_r b + q_o c + q_4 d + q_hl a^5 + q_hr b^5 + q_h4 d^5 + q_c) * q_arith_eval
        terms.push_back({&pk.coeffs[Q_M], g.a * g.b * qa});
        terms.push_back({&pk.coeffs[Q_L], g.a * qa});
        terms.push_back({&pk.coeffs[Q_R], g.b * qa});
        terms.push_back({&pk.coeffs[Q_O], g.c * qa});
        terms.push_back({&pk.coeffs[Q_4], g.d * qa});
        terms.push_back({&pk.coeffs[Q_HL], g.a.pow_u64(5) * qa});
        terms.push_back({&pk.coeffs[Q_HR], g.b.pow_u64(5) * qa});
        terms.push_back({&pk.coeffs[Q_H4], g.d.pow_u64(5) * qa});
        terms.push_back({&pk.coeffs[Q_C], qa});
        terms.push_back({&pk.coeffs[Q_RANGE], range_constraints(ch.range_sep, g)});
        terms.push_back({&pk.coeffs[Q_LOGIC], logic_constraints(ch.logic_sep, g)});
        terms.push_back({&pk.coeffs[Q_FIXED], fbsm_constraints(ch.fixed_sep, g)});
        terms.push_back({&pk.coeffs[Q_VAR], curve_add_constraints(ch.var_sep, g)});
        // lookup.rs:155-215
        Fr lsep_sq = ch.lookup_sep.square(), lsep_cu = ch.lookup_sep * lsep_sq;
        Fr opd = ch.delta + Fr::one(), eopd = ch.epsilon * opd;
        terms.push_back({&pk.coeffs[Q_LOOKUP], (lc4(g.a, g.b, g.c, g.d, ch.zeta) - e[E_F]) * ch.lookup_sep});
        {
            Fr b0 = ch.epsilon + e[E_F];
            Fr b1 = eopd + e[E_TABLE] + ch.delta * e[E_TABLENEXT];
            Fr b2 = l1_eval * lsep_cu;
            terms.push_back({&z2_poly, opd * b0 * b1 * lsep_sq + b2});
            Fr c0 = -e[E_Z2NEXT] * lsep_sq;
            Fr c1 = eopd + e[E_H2] + ch.delta * e[E_H1NEXT];
            terms.push_back({&h1_poly, c0 * c1});
        }
        // permutation.rs:156-296
        {
            Fr beta_z = ch.beta * ch.z;
            Fr a0 = g.a + beta_z + ch.gamma;
            Fr a1 = g.b + K_const(1) * beta_z + ch.gamma;
            Fr a2 = g.c + K_const(2) * beta_z + ch.gamma;
            Fr a3 = g.d + K_const(3) * beta_z + ch.gamma;
            Fr a = a0 * a1 * a2 * a3 * ch.alpha;
            terms.push_back({&z_poly, a});
            Fr b0 = g.a + ch.beta * e[E_LSIG] + ch.gamma;
            Fr b1 = g.b + ch.beta * e[E_RSIG] + ch.gamma;
            Fr b2 = g.c + ch.beta * e[E_OSIG] + ch.gamma;
            Fr b = b0 * b1 * b2 * (ch.beta * e[E_PERM]) * ch.alpha;
            terms.push_back({&pk.coeffs[SIG_4], -b});
            terms.push_back({&z_poly, l1_eval * ch.alpha.square()});
        }
        // quotient term: -(t_1 + z^n t_2 + ... + z^7n t_8) * Z_H(z)
        Fr zp = Fr::one();
        for (int k = 0; k < 8; k++) {
            terms.push_back({&t_i[k], -(zp * vanishing)});
            zp = zp * z_to_n;
        }
        for (auto& t : terms) {
            const std::vector<Fr>& p = *t.first;
            const Fr s = t.second;
#pragma omp parallel for schedule(static) if (n >= 4096)
            for (long i = 0; i < (long)p.size(); i++) lin[i] = lin[i] + p[i] * s;
        }
    }

    // evaluations into the transcript (prover.rs:532-572); table_eval / table_next_eval are not appended
    tr.append_fr("a_eval", e[E_A]);
    tr.append_fr("b_eval", e[E_B]);
    tr.append_fr("c_eval", e[E_C]);
    tr.append_fr("d_eval", e[E_D]);
    tr.append_fr("left_sig_eval", e[E_LSIG]);
    tr.append_fr("right_sig_eval", e[E_RSIG]);
    tr.append_fr("out_sig_eval", e[E_OSIG]);
    tr.append_fr("perm_eval", e[E_PERM]);
    tr.append_fr("f_eval", e[E_F]);
    tr.append_fr("q_lookup_eval", e[E_QLOOKUP]);
    tr.append_fr("lookup_perm_eval", e[E_Z2NEXT]);
    tr.append_fr("h_1_eval", e[E_H1]);
    tr.append_fr("h_1_next_eval", e[E_H1NEXT]);
    tr.append_fr("h_2_eval", e[E_H2]);
    {
        static const char* cl[10] = {"q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "q_hl_eval",
                                     "q_hr_eval", "q_h4_eval", "a_next_eval", "b_next_eval", "d_next_eval"};
        for (int k = 0; k < 10; k++) tr.append_fr(cl[k], e[E_QARITH + k]);
    }

    clk.lap("5b linearisation");
    // 6. openings
    ch.aw = tr.challenge_scalar("aggregate_witness");
    proof.comm[C_AW] = kzg_open(srs,
                                {&lin, &pk.coeffs[SIG_L], &pk.coeffs[SIG_R], &pk.coeffs[SIG_O], &f_poly, &h2_poly, &table_poly,
                                 &w_poly[0], &w_poly[1], &w_poly[2], &w_poly[3]},
                                ch.z, ch.aw);
    ch.saw = tr.challenge_scalar("aggregate_witness");
    proof.comm[C_SAW] = kzg_open(srs, {&z_poly, &w_poly[0], &w_poly[1], &w_poly[3], &h1_poly, &z2_poly, &table_poly}, zs, ch.saw);
    clk.lap("6 openings");
    if (ch_out) *ch_out = ch;
    return proof;
}

}  // namespace zpo
