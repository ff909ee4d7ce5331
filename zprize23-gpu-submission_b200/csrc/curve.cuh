// BLS12-381 G1 in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; ZZ = 0 is infinity).
// Replaces the reference's sppark `xyzz_t` / `jacobian_t`
// ("Prize 1B/plonk-core/lib/PLONK/utils/zkp/cuda/ec/xyzz_t.hpp":455-535, "…/ec/jacobian_t.hpp").
// Formulas: EFD "madd-2008-s" (mixed, 8M+2S), "add-2008-s" (12M+2S), "mdbl-2008-s-1"/"dbl-2008-s-1".
// All special cases (infinity, P == Q, P == -Q) are handled so the group law is exact on every input.
#pragma once
#include "field.cuh"

namespace zp {

struct affine_t {  // FFI layout: x || y, Montgomery, no infinity flag ("…/plonk-core/src/lib.rs":231-235)
    fq_t x, y;
};

struct xyzz_t {
    fq_t X, Y, ZZ, ZZZ;

    ZP_HD static xyzz_t infinity() {
        xyzz_t r;
        r.X = fq_t::zero();
        r.Y = fq_t::zero();
        r.ZZ = fq_t::zero();
        r.ZZZ = fq_t::zero();
        return r;
    }
    ZP_HD bool is_inf() const { return ZZ.is_zero(); }

    ZP_HD static xyzz_t from_affine(const affine_t& p) {
        xyzz_t r;
        r.X = p.x;
        r.Y = p.y;
        r.ZZ = fq_t::one();
        r.ZZZ = fq_t::one();
        return r;
    }

    // this = 2 * (x, y)
    ZP_HD void set_double_affine(const fq_t& x, const fq_t& y) {
        fq_t U = y.dbl();
        fq_t V = U.sqr();
        fq_t W = U * V;
        fq_t S = x * V;
        fq_t xx = x.sqr();
        fq_t M = xx.dbl() + xx;
        X = M.sqr() - S.dbl();
        Y = M * (S - X) - W * y;
        ZZ = V;
        ZZZ = W;
    }

    // this += (x, y)   (affine point assumed finite)
    ZP_HD void add_affine(const fq_t& x, const fq_t& y) {
        if (is_inf()) {
            X = x;
            Y = y;
            ZZ = fq_t::one();
            ZZZ = fq_t::one();
            return;
        }
        fq_t U2 = x * ZZ;
        fq_t S2 = y * ZZZ;
        fq_t P = U2 - X;
        fq_t R = S2 - Y;
        if (P.is_zero()) {
            if (R.is_zero()) {
                set_double_affine(x, y);
            } else {
                *this = infinity();
            }
            return;
        }
        fq_t PP = P.sqr();
        fq_t PPP = P * PP;
        fq_t Q = X * PP;
        fq_t X3 = R.sqr() - PPP - Q.dbl();
        fq_t Y3 = R * (Q - X3) - Y * PPP;
        X = X3;
        Y = Y3;
        ZZ = ZZ * PP;
        ZZZ = ZZZ * PPP;
    }

    ZP_HD void dbl_inplace() {
        if (is_inf()) return;
        fq_t U = Y.dbl();
        fq_t V = U.sqr();
        fq_t W = U * V;
        fq_t S = X * V;
        fq_t xx = X.sqr();
        fq_t M = xx.dbl() + xx;
        fq_t X3 = M.sqr() - S.dbl();
        fq_t Y3 = M * (S - X3) - W * Y;
        X = X3;
        Y = Y3;
        ZZ = V * ZZ;
        ZZZ = W * ZZZ;
    }

    // this += o
    ZP_HD void add(const xyzz_t& o) {
        if (o.is_inf()) return;
        if (is_inf()) {
            *this = o;
            return;
        }
        fq_t U1 = X * o.ZZ;
        fq_t U2 = o.X * ZZ;
        fq_t S1 = Y * o.ZZZ;
        fq_t S2 = o.Y * ZZZ;
        fq_t P = U2 - U1;
        fq_t R = S2 - S1;
        if (P.is_zero()) {
            if (R.is_zero()) {
                dbl_inplace();
            } else {
                *this = infinity();
            }
            return;
        }
        fq_t PP = P.sqr();
        fq_t PPP = P * PP;
        fq_t Q = U1 * PP;
        fq_t X3 = R.sqr() - PPP - Q.dbl();
        fq_t Y3 = R * (Q - X3) - S1 * PPP;
        X = X3;
        Y = Y3;
        ZZ = ZZ * o.ZZ * PP;
        ZZZ = ZZZ * o.ZZZ * PPP;
    }
};

}  // namespace zp
