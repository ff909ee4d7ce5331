// ORACLE — TEST INFRASTRUCTURE ONLY (see zp_field.hpp header).
//
// CPU restatement of the reference verifier `Proof::verify`
// ("Prize 1B/plonk-core/src/proof_system/proof.rs":123-443, compute_r0 :444-503,
// compute_linearisation_commitment :505-640, compute_first_lagrange_evaluation :647-658,
// compute_barycentric_eval :660-701; permutation VK "…/proof_system/permutation.rs":325-385;
// lookup VK "…/widget/lookup.rs":236-295; arithmetic VK "…/widget/arithmetic.rs":143-199).
//
// The two `PC::check` pairing equations  e(C - v*G, H) = e(W, tau*H - point*H)  are evaluated in G1
// with the KNOWN trapdoor of the synthetic SRS:  C - v*G == (tau - point) * W, which is equivalent for
// G1 inputs (SURVEY §8c pin (1)).  tests/test_oracle_vs_ref.py additionally re-checks the same
// equation with the reference's vendored blst pairing when oracle/_ref is built.
#pragma once
#include "zp_prover.hpp"

namespace zpo {

struct VerifierKeyO {
    int logn;
    size_t n;
    G1Affine pk_comm[NUM_PK_POLYS];
    G1Affine table_comm[4];
};

static inline VerifierKeyO make_verifier_key(const ProverKeyO& pk, const std::vector<G1Affine>& srs) {
    VerifierKeyO vk;
    vk.logn = pk.logn;
    vk.n = pk.n;
    Domain dom(pk.logn);
    for (int s = 0; s < NUM_PK_POLYS; s++) vk.pk_comm[s] = kzg_commit(srs, pk.coeffs[s]);
    for (int c = 0; c < 4; c++) vk.table_comm[c] = kzg_commit(srs, dom.ifft(pk.table[c]));
    return vk;
}

struct VerifyTrace {
    bool aw_ok, saw_ok;
    Challenges ch;
    G1Affine aw_C, saw_C;  // aggregated commitments (for the pairing re-check)
    Fr aw_v, saw_v;
};

static inline bool verify(const VerifierKeyO& vk, const ProofO& proof, const std::vector<std::pair<uint64_t, Fr>>& pi,
                          const std::string& label, const Fr& tau, VerifyTrace* trace = nullptr) {
    ensure_init();
    Domain dom(vk.logn);
    const size_t n = vk.n;
    const Fr* e = proof.eval;
    Challenges ch;
    Transcript tr(label);
    tr.append_pi("pi", pi);
    tr.append_g1("w_l", proof.comm[C_A]);
    tr.append_g1("w_r", proof.comm[C_B]);
    tr.append_g1("w_o", proof.comm[C_C]);
    tr.append_g1("w_4", proof.comm[C_D]);
    ch.zeta = tr.challenge_scalar("zeta");
    tr.append_fr("zeta", ch.zeta);
    tr.append_g1("f", proof.comm[C_F]);
    tr.append_g1("h1", proof.comm[C_H1]);
    tr.append_g1("h2", proof.comm[C_H2]);
    ch.beta = tr.challenge_scalar("beta");
    tr.append_fr("beta", ch.beta);
    ch.gamma = tr.challenge_scalar("gamma");
    tr.append_fr("gamma", ch.gamma);
    ch.delta = tr.challenge_scalar("delta");
    tr.append_fr("delta", ch.delta);
    ch.epsilon = tr.challenge_scalar("epsilon");
    tr.append_fr("epsilon", ch.epsilon);
    tr.append_g1("z", proof.comm[C_Z]);
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
    static const char* tl[8] = {"t_1", "t_2", "t_3", "t_4", "t_5", "t_6", "t_7", "t_8"};
    for (int k = 0; k < 8; k++) tr.append_g1(tl[k], proof.comm[C_T1 + k]);
    ch.z = tr.challenge_scalar("z");
    tr.append_fr("z", ch.z);

    Fr z_h_eval = dom.evaluate_vanishing(ch.z);
    Fr l1_eval = z_h_eval * (Fr::from_u64(n) * (ch.z - Fr::one())).inverse();

    // compute_r0 (proof.rs:444-503)
    Fr pi_eval = Fr::zero();
    {
        Fr numerator = z_h_eval * dom.n_inv;
        for (auto& p : pi) {
            Fr den = (dom.omega_inv.pow_u64(p.first) * ch.z) - Fr::one();
            pi_eval += den.inverse() * p.second;
        }
        pi_eval = pi_eval * numerator;
    }
    Fr alpha_sq = ch.alpha.square();
    Fr lsep_sq = ch.lookup_sep.square(), lsep_cu = lsep_sq * ch.lookup_sep;
    Fr r0;
    {
        Fr b0 = e[E_A] + ch.beta * e[E_LSIG] + ch.gamma;
        Fr b1 = e[E_B] + ch.beta * e[E_RSIG] + ch.gamma;
        Fr b2 = e[E_C] + ch.beta * e[E_OSIG] + ch.gamma;
        Fr b3 = (e[E_D] + ch.gamma) * e[E_PERM] * ch.alpha;
        Fr b = b0 * b1 * b2 * b3;
        Fr c = l1_eval * alpha_sq;
        Fr eopd = ch.epsilon * (Fr::one() + ch.delta);
        Fr d0 = lsep_sq * e[E_Z2NEXT];
        Fr d1 = eopd + ch.delta * e[E_H2];
        Fr d2 = eopd + e[E_H2] + ch.delta * e[E_H1NEXT];
        Fr d = d0 * d1 * d2;
        Fr ee = lsep_cu * l1_eval;
        r0 = pi_eval - b - c - d - ee;
    }

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
    static const char* cl[10] = {"q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "q_hl_eval",
                                 "q_hr_eval", "q_h4_eval", "a_next_eval", "b_next_eval", "d_next_eval"};
    for (int k = 0; k < 10; k++) tr.append_fr(cl[k], e[E_QARITH + k]);

    // linearisation commitment (proof.rs:505-640)
    std::vector<Fr> scalars;
    std::vector<G1Affine> points;
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
        Fr qa = e[E_QARITH];
        auto push = [&](const Fr& s, const G1Affine& p) {
            scalars.push_back(s);
            points.push_back(p);
        };
        push(g.a * g.b * qa, vk.pk_comm[Q_M]);
        push(g.a * qa, vk.pk_comm[Q_L]);
        push(g.b * qa, vk.pk_comm[Q_R]);
        push(g.d * qa, vk.pk_comm[Q_4]);
        push(g.c * qa, vk.pk_comm[Q_O]);
        push(g.a.pow_u64(5) * qa, vk.pk_comm[Q_HL]);
        push(g.b.pow_u64(5) * qa, vk.pk_comm[Q_HR]);
        push(g.d.pow_u64(5) * qa, vk.pk_comm[Q_H4]);
        push(qa, vk.pk_comm[Q_C]);
        push(range_constraints(ch.range_sep, g), vk.pk_comm[Q_RANGE]);
        push(logic_constraints(ch.logic_sep, g), vk.pk_comm[Q_LOGIC]);
        push(fbsm_constraints(ch.fixed_sep, g), vk.pk_comm[Q_FIXED]);
        push(curve_add_constraints(ch.var_sep, g), vk.pk_comm[Q_VAR]);
        // lookup VK
        Fr opd = Fr::one() + ch.delta, eopd = ch.epsilon * opd;
        push((lc4(g.a, g.b, g.c, g.d, ch.zeta) - e[E_F]) * ch.lookup_sep, vk.pk_comm[Q_LOOKUP]);
        {
            Fr b0 = ch.epsilon + e[E_F];
            Fr b1 = eopd + e[E_TABLE] + ch.delta * e[E_TABLENEXT];
            Fr b2 = l1_eval * lsep_cu;
            push(opd * b0 * b1 * lsep_sq + b2, proof.comm[C_Z2]);
            Fr c0 = -e[E_Z2NEXT] * lsep_sq;
            Fr c1 = eopd + e[E_H2] + ch.delta * e[E_H1NEXT];
            push(c0 * c1, proof.comm[C_H1]);
        }
        // permutation VK
        {
            Fr beta_z = ch.beta * ch.z;
            Fr q0 = g.a + beta_z + ch.gamma;
            Fr q1 = g.b + ch.beta * K_const(1) * ch.z + ch.gamma;
            Fr q2 = g.c + ch.beta * K_const(2) * ch.z + ch.gamma;
            Fr q3 = (g.d + ch.beta * K_const(3) * ch.z + ch.gamma) * ch.alpha;
            Fr x = q0 * q1 * q2 * q3;
            Fr r = l1_eval * alpha_sq;
            push(x + r, proof.comm[C_Z]);
            Fr y0 = g.a + ch.beta * e[E_LSIG] + ch.gamma;
            Fr y1 = g.b + ch.beta * e[E_RSIG] + ch.gamma;
            Fr y2 = g.c + ch.beta * e[E_OSIG] + ch.gamma;
            Fr y3 = ch.beta * e[E_PERM] * ch.alpha;
            push(-(y0 * y1 * y2 * y3), vk.pk_comm[SIG_4]);
        }
        Fr z_to_n = z_h_eval + Fr::one();
        Fr ts = -z_h_eval;
        for (int k = 0; k < 8; k++) {
            push(ts, proof.comm[C_T1 + k]);
            ts = ts * z_to_n;
        }
    }
    G1 lin_comm = g1_msm(points.data(), scalars.data(), points.size());
    Fr zeta_sq = ch.zeta.square();
    Fr tsc[4] = {Fr::one(), ch.zeta, zeta_sq, zeta_sq * ch.zeta};
    G1 table_comm = g1_msm(vk.table_comm, tsc, 4);

    ch.aw = tr.challenge_scalar("aggregate_witness");
    ch.saw = tr.challenge_scalar("aggregate_witness");

    G1 gen = G1::from_affine(g1_generator());
    auto check = [&](const std::vector<G1>& comms, const std::vector<Fr>& evals, const Fr& point, const Fr& chal,
                     const G1Affine& W, G1Affine* Cout, Fr* vout) {
        G1 C = G1::infinity();
        Fr v = Fr::zero(), cj = Fr::one();
        for (size_t j = 0; j < comms.size(); j++) {
            C = C.add(comms[j].mul(cj));
            v += evals[j] * cj;
            cj = cj * chal;
        }
        if (Cout) *Cout = C.to_affine();
        if (vout) *vout = v;
        G1 lhs = C.add(gen.mul(v).neg());
        G1 rhs = G1::from_affine(W).mul(tau - point);
        return lhs.to_affine() == rhs.to_affine();
    };
    auto A = [&](int c) { return G1::from_affine(proof.comm[c]); };
    std::vector<G1> aw_comms = {lin_comm, G1::from_affine(vk.pk_comm[SIG_L]), G1::from_affine(vk.pk_comm[SIG_R]),
                                G1::from_affine(vk.pk_comm[SIG_O]), A(C_F), A(C_H2), table_comm, A(C_A), A(C_B), A(C_C), A(C_D)};
    std::vector<Fr> aw_evals = {-r0, e[E_LSIG], e[E_RSIG], e[E_OSIG], e[E_F], e[E_H2], e[E_TABLE], e[E_A], e[E_B], e[E_C], e[E_D]};
    std::vector<G1> saw_comms = {A(C_Z), A(C_A), A(C_B), A(C_D), A(C_H1), A(C_Z2), table_comm};
    std::vector<Fr> saw_evals = {e[E_PERM], e[E_ANEXT], e[E_BNEXT], e[E_DNEXT], e[E_H1NEXT], e[E_Z2NEXT], e[E_TABLENEXT]};
    VerifyTrace t;
    t.ch = ch;
    t.aw_ok = check(aw_comms, aw_evals, ch.z, ch.aw, proof.comm[C_AW], &t.aw_C, &t.aw_v);
    t.saw_ok = check(saw_comms, saw_evals, ch.z * dom.omega, ch.saw, proof.comm[C_SAW], &t.saw_C, &t.saw_v);
    if (trace) *trace = t;
    return t.aw_ok && t.saw_ok;
}

}  // namespace zpo
