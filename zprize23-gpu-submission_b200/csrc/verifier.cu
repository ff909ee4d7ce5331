// Product-side verifier: `Proof::verify` of the reference ("Prize 1B/plonk-core/src/proof_system/proof.rs":123-443;
// compute_r0 :444-503, linearisation commitment :505-640, L_1 :647-658; KZG check = ark-poly-commit 0.3 `KZG10::check`,
// called at proof.rs:414-441 and from `Verifier::verify`, verifier.rs:106-125) with REAL pairings (pairing.hpp), plus a
// batch form that folds any number of proofs into one two-pairing product.  Host code: the reference verifies on the CPU
// as well; it is ~40 group operations and two pairing products per proof.
//
// KZG opening check for commitment C, value v, point z, witness W, with H the G2 generator and [tau]H from the SRS:
//     e(C - v G, H) = e(W, [tau]H - z H)   <=>   e(C - v G + z W, H) * e(-W, [tau]H) = 1
// (the right-hand form needs no G2 arithmetic per proof).
#include "prover.cuh"
#include "gates.cuh"
#include "pairing.hpp"

namespace zp {

using host::Fq;
using host::Fr;
using host::G1;

namespace {

struct VerifierCtx {
    int logn = 0;
    size_t n = 0;
    G1 pk_comm[PK_COUNT];
    G1 table_comm[4];
    host::G2Affine beta_h;  // [tau] H
    std::string label = "Merkle tree";
};

thread_local std::string v_err;
template <class F>
int vguard(F&& f) {
    try {
        f();
        v_err.clear();
        return 0;
    } catch (const std::exception& e) {
        v_err = e.what();
        return -1;
    }
}

G1 g1_from_words(const uint64_t* w, bool check_curve) {
    Fq x, y;
    memcpy(x.v, w, 48);
    memcpy(y.v, w + 6, 48);
    if (x.is_zero() && y == Fq::one()) return G1::infinity();  // FFI encoding of the identity (point.cu:30-34)
    if (check_curve && !(y.sqr() == x.sqr() * x + Fq::from_u64(4))) throw std::runtime_error("verifier: point not on the curve");
    return G1::from_affine(x, y);
}
G1 g1_neg(const G1& p) {
    G1 r = p;
    r.Y = r.Y.neg();
    return r;
}
G1 g1_mul(const G1& p, const Fr& s) {
    uint64_t k[4];
    s.to_canonical(k);
    G1 acc = G1::infinity();
    bool started = false;
    for (int i = 254; i >= 0; i--) {
        if (started) acc.dbl_inplace();
        if ((k[i >> 6] >> (i & 63)) & 1) {
            acc.add(p);
            started = true;
        }
    }
    return acc;
}
G1 g1_generator() {
    static const uint64_t gx[6] = {0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL, 0xa14e3a3f171bac58ULL,
                                   0xc3688c4f9774b905ULL, 0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL};
    static const uint64_t gy[6] = {0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL, 0x00db18cb2c04b3edULL,
                                   0xfcf5e095d5d00af6ULL, 0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL};
    return G1::from_affine(Fq::from_canonical(gx), Fq::from_canonical(gy));
}
host::G2Affine g2_from_words(const uint64_t* w) {
    host::G2Affine q;
    memcpy(q.x.c0.v, w, 48);
    memcpy(q.x.c1.v, w + 6, 48);
    memcpy(q.y.c0.v, w + 12, 48);
    memcpy(q.y.c1.v, w + 18, 48);
    q.inf = q.x.is_zero() && q.y.is_zero();
    if (!host::g2_on_curve(q)) throw std::runtime_error("verifier: G2 point not on the twist");
    return q;
}

// One opening equation reduced to its two G1 sides:  e(lhs, H) * e(-W, [tau]H) = 1
struct Opening {
    G1 lhs;  // C - v G + point W
    G1 w;
};
struct ProofOpenings {
    Opening aw, saw;
};

// Everything of Proof::verify up to the two KZG checks: transcript replay, r0, linearisation commitment, aggregation.
ProofOpenings reduce_proof(const VerifierCtx& vk, const ProofC& proof, const std::vector<std::pair<uint64_t, Fr>>& pi) {
    const uint64_t* cw = reinterpret_cast<const uint64_t*>(&proof.a_comm);
    enum { C_A = 0, C_B, C_C, C_D, C_Z, C_F, C_H1, C_H2, C_Z2, C_T1, C_AW = 17, C_SAW = 18 };
    G1 comm[19];
    Fq cx[19], cy[19];
    bool cinf[19];
    for (int i = 0; i < 19; i++) {
        comm[i] = g1_from_words(cw + 12 * i, true);
        comm[i].to_affine(cx[i], cy[i], cinf[i]);
    }
    enum { E_A = 0, E_B, E_C, E_D, E_LSIG, E_RSIG, E_OSIG, E_PERM, E_QLOOKUP, E_Z2NEXT, E_H1, E_H1NEXT, E_H2, E_F, E_TABLE,
           E_TABLENEXT, E_QARITH, E_QC, E_QL, E_QR, E_QHL, E_QHR, E_QH4, E_ANEXT, E_BNEXT, E_DNEXT, NUM_E };
    Fr e[NUM_E];
    const uint64_t* ew = reinterpret_cast<const uint64_t*>(&proof.evaluations);
    for (int i = 0; i < NUM_E; i++) memcpy(e[i].v, ew + 4 * i, 32);

    MerlinTranscript tr(vk.label);
    tr.append_public_inputs("pi", pi);
    auto app = [&](const char* l, int c) { tr.append_point(l, cx[c], cy[c], cinf[c]); };
    app("w_l", C_A); app("w_r", C_B); app("w_o", C_C); app("w_4", C_D);
    Fr zeta = tr.challenge_scalar("zeta");
    tr.append_scalar("zeta", zeta);
    app("f", C_F); app("h1", C_H1); app("h2", C_H2);
    Fr beta = tr.challenge_scalar("beta");
    tr.append_scalar("beta", beta);
    Fr gamma = tr.challenge_scalar("gamma");
    tr.append_scalar("gamma", gamma);
    Fr delta = tr.challenge_scalar("delta");
    tr.append_scalar("delta", delta);
    Fr epsilon = tr.challenge_scalar("epsilon");
    tr.append_scalar("epsilon", epsilon);
    app("z", C_Z);
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
    static const char* tl[8] = {"t_1", "t_2", "t_3", "t_4", "t_5", "t_6", "t_7", "t_8"};
    for (int k = 0; k < 8; k++) app(tl[k], C_T1 + k);
    Fr z = tr.challenge_scalar("z");
    tr.append_scalar("z", z);

    // domain constants
    Fr omega;
    {
        fr_t w = fr_two_adic_root_host();
        for (int i = vk.logn; i < 32; i++) w = w.sqr();
        omega = host::to_host(w);
    }
    const Fr one = Fr::one();
    Fr n_fr = Fr::from_u64(vk.n);
    Fr z_h = z.pow_u64(vk.n) - one;
    Fr l1 = z_h * (n_fr * (z - one)).inverse();  // proof.rs:647-658
    // PI(z) by the barycentric formula over the non-zero inputs (proof.rs:660-701)
    Fr pi_eval = Fr::zero();
    {
        Fr omega_inv = omega.inverse();
        for (auto& p : pi) pi_eval = pi_eval + (omega_inv.pow_u64(p.first) * z - one).inverse() * p.second;
        pi_eval = pi_eval * z_h * n_fr.inverse();
    }
    Fr alpha_sq = alpha.sqr(), lsep_sq = lookup_sep.sqr(), lsep_cu = lsep_sq * lookup_sep;
    Fr opd = one + delta, eopd = epsilon * opd;
    // compute_r0 (proof.rs:444-503)
    Fr r0;
    {
        Fr b = (e[E_A] + beta * e[E_LSIG] + gamma) * (e[E_B] + beta * e[E_RSIG] + gamma) * (e[E_C] + beta * e[E_OSIG] + gamma) *
               ((e[E_D] + gamma) * e[E_PERM] * alpha);
        Fr c = l1 * alpha_sq;
        Fr d = (lsep_sq * e[E_Z2NEXT]) * (eopd + delta * e[E_H2]) * (eopd + e[E_H2] + delta * e[E_H1NEXT]);
        r0 = pi_eval - b - c - d - lsep_cu * l1;
    }
    static const char* en[NUM_E] = {"a_eval", "b_eval", "c_eval", "d_eval", "left_sig_eval", "right_sig_eval", "out_sig_eval",
                                    "perm_eval", "q_lookup_eval", "lookup_perm_eval", "h_1_eval", "h_1_next_eval", "h_2_eval",
                                    "f_eval", "", "", "q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "q_hl_eval", "q_hr_eval",
                                    "q_h4_eval", "a_next_eval", "b_next_eval", "d_next_eval"};
    static const int order[] = {E_A, E_B, E_C, E_D, E_LSIG, E_RSIG, E_OSIG, E_PERM, E_F, E_QLOOKUP, E_Z2NEXT, E_H1, E_H1NEXT, E_H2,
                                E_QARITH, E_QC, E_QL, E_QR, E_QHL, E_QHR, E_QH4, E_ANEXT, E_BNEXT, E_DNEXT};
    for (int idx : order) tr.append_scalar(en[idx], e[idx]);

    // linearisation commitment (proof.rs:505-640)
    G1 lin = G1::infinity();
    auto acc = [&](const Fr& s, const G1& p) {
        if (!p.is_inf() && !s.is_zero()) lin.add(g1_mul(p, s));
    };
    {
        GateVals<Fr> g;
        g.a = e[E_A]; g.b = e[E_B]; g.c = e[E_C]; g.d = e[E_D];
        g.a_next = e[E_ANEXT]; g.b_next = e[E_BNEXT]; g.d_next = e[E_DNEXT];
        g.q_l = e[E_QL]; g.q_r = e[E_QR]; g.q_c = e[E_QC];
        auto pow5 = [](const Fr& x) { Fr s = x.sqr(); return s.sqr() * x; };
        static const uint64_t JUBJUB_D[4] = {3049539848285517488ULL, 18189135023605205683ULL, 8793554888777148625ULL,
                                             6339087681201251886ULL};  // edwards.cu:20-31
        Fr coeff_d;
        memcpy(coeff_d.v, JUBJUB_D, 32);
        Fr qa = e[E_QARITH];
        acc(g.a * g.b * qa, vk.pk_comm[PK_QM]);  // arithmetic.rs:143-199
        acc(g.a * qa, vk.pk_comm[PK_QL]);
        acc(g.b * qa, vk.pk_comm[PK_QR]);
        acc(g.c * qa, vk.pk_comm[PK_QO]);
        acc(g.d * qa, vk.pk_comm[PK_Q4]);
        acc(pow5(g.a) * qa, vk.pk_comm[PK_QHL]);
        acc(pow5(g.b) * qa, vk.pk_comm[PK_QHR]);
        acc(pow5(g.d) * qa, vk.pk_comm[PK_QH4]);
        acc(qa, vk.pk_comm[PK_QC]);
        acc(range_constraints(range_sep, g), vk.pk_comm[PK_RANGE]);
        acc(logic_constraints(logic_sep, g), vk.pk_comm[PK_LOGIC]);
        acc(fbsm_constraints(fixed_sep, g, coeff_d), vk.pk_comm[PK_FIXED]);
        acc(curve_add_constraints(var_sep, g, coeff_d), vk.pk_comm[PK_VAR]);
        // lookup (widget/lookup.rs:236-295)
        acc((lc4(g.a, g.b, g.c, g.d, zeta) - e[E_F]) * lookup_sep, vk.pk_comm[PK_QLOOKUP]);
        acc(opd * (epsilon + e[E_F]) * (eopd + e[E_TABLE] + delta * e[E_TABLENEXT]) * lsep_sq + l1 * lsep_cu, comm[C_Z2]);
        acc((Fr::zero() - e[E_Z2NEXT] * lsep_sq) * (eopd + e[E_H2] + delta * e[E_H1NEXT]), comm[C_H1]);
        // permutation (proof_system/permutation.rs:325-385)
        Fr bz = beta * z;
        Fr k1 = Fr::from_u64(7), k2 = Fr::from_u64(13), k3 = Fr::from_u64(17);
        acc((g.a + bz + gamma) * (g.b + k1 * bz + gamma) * (g.c + k2 * bz + gamma) * ((g.d + k3 * bz + gamma) * alpha) + l1 * alpha_sq,
            comm[C_Z]);
        acc(Fr::zero() - (g.a + beta * e[E_LSIG] + gamma) * (g.b + beta * e[E_RSIG] + gamma) * (g.c + beta * e[E_OSIG] + gamma) *
                             (beta * e[E_PERM] * alpha),
            vk.pk_comm[PK_SIG4]);
        Fr z_to_n = z_h + one, ts = Fr::zero() - z_h;
        for (int k = 0; k < 8; k++) {
            acc(ts, comm[C_T1 + k]);
            ts = ts * z_to_n;
        }
    }
    // table commitment t_1 + zeta t_2 + zeta^2 t_3 + zeta^3 t_4 (proof.rs:333-342)
    G1 table = G1::infinity();
    {
        Fr p = one;
        for (int c = 0; c < 4; c++) {
            if (!vk.table_comm[c].is_inf()) table.add(g1_mul(vk.table_comm[c], p));
            p = p * zeta;
        }
    }
    Fr aw = tr.challenge_scalar("aggregate_witness");
    Fr saw = tr.challenge_scalar("aggregate_witness");
    const G1 gen = g1_generator();
    auto aggregate = [&](const std::vector<const G1*>& cs, const std::vector<Fr>& vs, const Fr& point, const Fr& chal, const G1& W) {
        G1 C = G1::infinity();
        Fr v = Fr::zero(), cj = one;
        for (size_t j = 0; j < cs.size(); j++) {
            if (!cs[j]->is_inf()) C.add(g1_mul(*cs[j], cj));
            v = v + vs[j] * cj;
            cj = cj * chal;
        }
        Opening o;
        o.lhs = C;
        o.lhs.add(g1_neg(g1_mul(gen, v)));
        if (!W.is_inf()) o.lhs.add(g1_mul(W, point));
        o.w = W;
        return o;
    };
    ProofOpenings out;
    out.aw = aggregate({&lin, &vk.pk_comm[PK_SIGL], &vk.pk_comm[PK_SIGR], &vk.pk_comm[PK_SIGO], &comm[C_F], &comm[C_H2], &table,
                        &comm[C_A], &comm[C_B], &comm[C_C], &comm[C_D]},
                       {Fr::zero() - r0, e[E_LSIG], e[E_RSIG], e[E_OSIG], e[E_F], e[E_H2], e[E_TABLE], e[E_A], e[E_B], e[E_C], e[E_D]},
                       z, aw, comm[C_AW]);
    out.saw = aggregate({&comm[C_Z], &comm[C_A], &comm[C_B], &comm[C_D], &comm[C_H1], &comm[C_Z2], &table},
                        {e[E_PERM], e[E_ANEXT], e[E_BNEXT], e[E_DNEXT], e[E_H1NEXT], e[E_Z2NEXT], e[E_TABLENEXT]}, z * omega, saw,
                        comm[C_SAW]);
    return out;
}

// e(lhs, H) * e(-W, [tau]H) == 1
bool pairing_check(const VerifierCtx& vk, const G1& lhs, const G1& w) {
    std::vector<host::PairingInput> in(2);
    in[0].q = host::g2_generator();
    lhs.to_affine(in[0].px, in[0].py, in[0].p_inf);
    in[1].q = vk.beta_h;
    g1_neg(w).to_affine(in[1].px, in[1].py, in[1].p_inf);
    return host::pairing_product(in).is_one();
}

std::vector<std::pair<uint64_t, Fr>> pi_list(const uint64_t* pos, const uint64_t* vals, size_t cnt) {
    std::vector<std::pair<uint64_t, Fr>> pi;
    for (size_t i = 0; i < cnt; i++) {
        Fr v;
        memcpy(v.v, vals + 4 * i, 32);
        if (!v.is_zero()) pi.push_back({pos[i], v});  // PublicInputs keeps non-zero values only (pi.rs:55-62)
    }
    return pi;
}

inline VerifierCtx* V(zp_verifier* v) { return reinterpret_cast<VerifierCtx*>(v); }

}  // namespace
}  // namespace zp

using namespace zp;

extern "C" {

const char* zp_verifier_last_error(void) { return v_err.c_str(); }

zp_verifier* zp_verifier_create(uint64_t n, const uint64_t* commitments23, const uint64_t* beta_h) {
    VerifierCtx* c = nullptr;
    if (vguard([&] {
            int logn = ilog2((size_t)n);
            if (((uint64_t)1 << logn) != n || logn < 1 || logn > 32) throw std::runtime_error("zp_verifier_create: n must be a power of two");
            c = new VerifierCtx();
            c->logn = logn;
            c->n = (size_t)n;
            try {
                for (int i = 0; i < PK_COUNT; i++) c->pk_comm[i] = g1_from_words(commitments23 + 12 * i, true);
                for (int i = 0; i < 4; i++) c->table_comm[i] = g1_from_words(commitments23 + 12 * (PK_COUNT + i), true);
                c->beta_h = g2_from_words(beta_h);
            } catch (...) {
                delete c;
                c = nullptr;
                throw;
            }
        }))
        return nullptr;
    return reinterpret_cast<zp_verifier*>(c);
}
void zp_verifier_destroy(zp_verifier* v) { delete V(v); }
int zp_verifier_set_label(zp_verifier* v, const char* label) { return vguard([&] { V(v)->label = label; }); }

int zp_g2_mul_generator(const uint64_t* scalar, uint64_t* out24) {
    return vguard([&] {
        Fr s;
        memcpy(s.v, scalar, 32);
        host::G2Affine q = host::g2_mul(host::g2_generator(), s);
        if (q.inf) {
            memset(out24, 0, 192);
            return;
        }
        memcpy(out24, q.x.c0.v, 48);
        memcpy(out24 + 6, q.x.c1.v, 48);
        memcpy(out24 + 12, q.y.c0.v, 48);
        memcpy(out24 + 18, q.y.c1.v, 48);
    });
}

int zp_pairing_product(const uint64_t* g1_points, const uint64_t* g2_points, size_t count, uint64_t* out72, int* is_one) {
    return vguard([&] {
        std::vector<host::PairingInput> in(count);
        for (size_t i = 0; i < count; i++) {
            G1 p = g1_from_words(g1_points + 12 * i, true);
            p.to_affine(in[i].px, in[i].py, in[i].p_inf);
            in[i].q = g2_from_words(g2_points + 24 * i);
        }
        host::Fq12 r = host::pairing_product(in);
        if (out72) memcpy(out72, &r, 576);
        if (is_one) *is_one = r.is_one() ? 1 : 0;
    });
}

int zp_proof_verify(zp_verifier* v, const ProofC* proof, const uint64_t* pi_pos, const uint64_t* pi_vals, size_t n_pi, int* accepted,
                    int* detail) {
    if (accepted) *accepted = 0;
    if (detail) *detail = 0;
    return vguard([&] {
        const VerifierCtx& vk = *V(v);
        ProofOpenings o = reduce_proof(vk, *proof, pi_list(pi_pos, pi_vals, n_pi));
        bool a = pairing_check(vk, o.aw.lhs, o.aw.w), s = pairing_check(vk, o.saw.lhs, o.saw.w);
        if (detail) *detail = (a ? 1 : 0) | (s ? 2 : 0);
        if (accepted) *accepted = (a && s) ? 1 : 0;
    });
}

int zp_proof_verify_batch(zp_verifier* v, const ProofC* proofs, size_t count, const uint64_t* pi_pos, const uint64_t* pi_vals,
                          int* accepted) {
    if (accepted) *accepted = 0;
    return vguard([&] {
        const VerifierCtx& vk = *V(v);
        // random linear combination of all 2 * count opening equations; the coefficients come from a transcript over every
        // proof (deterministic, and no prover can choose a proof after seeing them)
        MerlinTranscript tr("zprize_b200 batch verify");
        tr.append_message("proofs", reinterpret_cast<const uint8_t*>(proofs), count * sizeof(ProofC));
        tr.append_message("pi", reinterpret_cast<const uint8_t*>(pi_vals), count * 32);
        G1 L = G1::infinity(), W = G1::infinity();
        for (size_t i = 0; i < count; i++) {
            ProofOpenings o = reduce_proof(vk, proofs[i], pi_list(pi_pos + i, pi_vals + 4 * i, 1));
            for (const Opening* op : {&o.aw, &o.saw}) {
                Fr r = tr.challenge_scalar("r");
                if (!op->lhs.is_inf()) L.add(g1_mul(op->lhs, r));
                if (!op->w.is_inf()) W.add(g1_mul(op->w, r));
            }
        }
        if (accepted) *accepted = pairing_check(vk, L, W) ? 1 : 0;
    });
}

}  // extern "C"
