// Gate-constraint formulas of the TurboPLONK arithmetisation, written once over a generic field type F
// (device `fr_t` inside the fused quotient kernel; host `host::Fr` for the linearisation scalars).
// They follow the ZK-Garage widgets term by term:
//   arithmetic   "Prize 1B/plonk-core/src/proof_system/widget/arithmetic.rs":61-79
//   range        "…/widget/range.rs":43-74
//   logic        "…/widget/logic.rs":60-133
//   fixed-base   "…/widget/ecc/fixed_base_scalar_mul.rs":82-156
//   curve add    "…/widget/ecc/curve_addition.rs":52-97
//   lookup       "…/widget/lookup.rs":98-152
//   permutation  "…/proof_system/permutation.rs":62-153
// (PNP's GPU twins: lib/PLONK/src/plonk_core/src/proof_system/widget/*.cu — not followed where they
// deviate from the Rust semantics)
#pragma once
#include "field.cuh"

namespace zp {

template <class F>
struct GateVals {
    F a, b, c, d, a_next, b_next, d_next, q_l, q_r, q_c;
};

// small constants in Montgomery form, built by repeated addition (no table needed on the device)
template <class F>
ZP_HD F small_const(uint32_t k) {
    F one = F::one(), r = F::zero(), p = one;
    // binary expansion of k
    while (k) {
        if (k & 1u) r = r + p;
        p = p + p;
        k >>= 1;
    }
    return r;
}

template <class F>
ZP_HD F delta4(const F& f, const F& one) {  // f (f-1)(f-2)(f-3)
    F two = one + one, three = two + one;
    return f * (f - one) * (f - two) * (f - three);
}

template <class F>
ZP_HD F range_constraints(const F& sep, const GateVals<F>& g) {
    F one = F::one();
    F kappa = sep.sqr(), kappa_sq = kappa.sqr(), kappa_cu = kappa_sq * kappa;
    F b1 = delta4(g.c - g.d.dbl().dbl(), one);
    F b2 = delta4(g.b - g.c.dbl().dbl(), one) * kappa;
    F b3 = delta4(g.a - g.b.dbl().dbl(), one) * kappa_sq;
    F b4 = delta4(g.d_next - g.a.dbl().dbl(), one) * kappa_cu;
    return (b1 + b2 + b3 + b4) * sep;
}

template <class F>
ZP_HD F delta_xor_and(const F& a, const F& b, const F& w, const F& c, const F& q_c) {
    F nine = small_const<F>(9), two = small_const<F>(2), three = small_const<F>(3), four = small_const<F>(4);
    F eighteen = small_const<F>(18), eighty_one = small_const<F>(81), eighty_three = small_const<F>(83);
    F Fv = w * (w * (four * w - eighteen * (a + b) + eighty_one) + eighteen * (a.sqr() + b.sqr()) - eighty_one * (a + b) +
                eighty_three);
    F E = three * (a + b + c) - (two * Fv);
    F B = q_c * ((nine * c) - three * (a + b));
    return B + E;
}

template <class F>
ZP_HD F logic_constraints(const F& sep, const GateVals<F>& g) {
    F one = F::one();
    F kappa = sep.sqr(), kappa_sq = kappa.sqr(), kappa_cu = kappa_sq * kappa, kappa_qu = kappa_cu * kappa;
    F a = g.a_next - g.a.dbl().dbl();
    F c0 = delta4(a, one);
    F b = g.b_next - g.b.dbl().dbl();
    F c1 = delta4(b, one) * kappa;
    F d = g.d_next - g.d.dbl().dbl();
    F c2 = delta4(d, one) * kappa_sq;
    F w = g.c;
    F c3 = (w - a * b) * kappa_cu;
    F c4 = delta_xor_and(a, b, w, d, g.q_c) * kappa_qu;
    return (c0 + c1 + c2 + c3 + c4) * sep;
}

// JubJub (ed-on-bls12-381): a = -1, d passed in (Montgomery literal in prover.cu,
// same value as "…/lib/PLONK/src/bls12_381/edwards.cu":20-31)
template <class F>
ZP_HD F fbsm_constraints(const F& sep, const GateVals<F>& g, const F& coeff_d) {
    F one = F::one();
    F kappa = sep.sqr(), kappa_sq = kappa.sqr(), kappa_cu = kappa_sq * kappa;
    F x_beta = g.q_l, y_beta = g.q_r;
    F acc_x = g.a, acc_x_next = g.a_next, acc_y = g.b, acc_y_next = g.b_next;
    F xy_alpha = g.c;
    F bit = g.d_next - g.d - g.d;
    F bit_consistency = bit * (bit - one) * (bit + one);
    F y_alpha = bit.sqr() * (y_beta - one) + one;
    F x_alpha = x_beta * bit;
    F xy_consistency = ((bit * g.q_c) - xy_alpha) * kappa;
    F x_3 = acc_x_next;
    F lhs = x_3 + (x_3 * xy_alpha * acc_x * acc_y * coeff_d);
    F rhs = (x_alpha * acc_y) + (y_alpha * acc_x);
    F x_acc = (lhs - rhs) * kappa_sq;
    F y_3 = acc_y_next;
    lhs = y_3 - (y_3 * xy_alpha * acc_x * acc_y * coeff_d);
    rhs = y_alpha * acc_y + x_alpha * acc_x;  // - COEFF_A * x_alpha * acc_x with COEFF_A = -1
    F y_acc = (lhs - rhs) * kappa_cu;
    return (bit_consistency + x_acc + y_acc + xy_consistency) * sep;
}

template <class F>
ZP_HD F curve_add_constraints(const F& sep, const GateVals<F>& g, const F& coeff_d) {
    F x_1 = g.a, x_3 = g.a_next, y_1 = g.b, y_3 = g.b_next, x_2 = g.c, y_2 = g.d, x1_y2 = g.d_next;
    F kappa = sep.sqr();
    F xy_consistency = x_1 * y_2 - x1_y2;
    F y1_x2 = y_1 * x_2, y1_y2 = y_1 * y_2, x1_x2 = x_1 * x_2;
    F x3_lhs = x1_y2 + y1_x2;
    F x3_rhs = x_3 + (x_3 * coeff_d * x1_y2 * y1_x2);
    F x3_consistency = (x3_lhs - x3_rhs) * kappa;
    F y3_lhs = y1_y2 + x1_x2;  // - COEFF_A * x1_x2 with COEFF_A = -1
    F y3_rhs = y_3 - y_3 * coeff_d * x1_y2 * y1_x2;
    F y3_consistency = (y3_lhs - y3_rhs) * kappa.sqr();
    return (xy_consistency + x3_consistency + y3_consistency) * sep;
}

template <class F>
ZP_HD F lc4(const F& a, const F& b, const F& c, const F& d, const F& ch) {  // util.rs:154-176
    return ((d * ch + c) * ch + b) * ch + a;
}

}  // namespace zp
