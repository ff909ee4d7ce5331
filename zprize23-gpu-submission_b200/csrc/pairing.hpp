// BLS12-381 optimal ate pairing on the host, for the product-side verifier (verifier.cu).
//
// The reference verifies with ark-ec 0.3 `Bls12::product_of_pairings` behind `KZG10::check`
// ("Prize 1B/plonk-core/src/proof_system/proof.rs":414-441; crate not vendored).  This file restates the textbook
// construction with formulas that can be checked by hand, and tests/test_verifier.py pins it against the reference's own
// vendored blst (`blst_miller_loop` + `blst_final_exp`, oracle/_ref/libref_blst.so) on the final GT bytes (blst's x-chain
// raises to 3 (q^12 - 1) / r, so the test compares the cube of our value):
//   tower    Fq2 = Fq[u]/(u^2 + 1),  Fq6 = Fq2[v]/(v^3 - xi),  xi = 1 + u,  Fq12 = Fq6[w]/(w^2 - v)   (so w^6 = xi)
//   twist    E'(Fq2): y^2 = x^3 + 4 xi  (M-type);  untwist  psi(x', y') = (x' w^-2, y' w^-3)
//   line     through T' with twist slope l' evaluated at P = (xP, yP), scaled by w^3 (an element of Fq4, erased by the
//            final exponentiation):   (l' xT' - yT')  +  (-l' xP) v  +  yP v w        -> coefficients 0, 1, 4 of Fq12
//   loop     f <- f^2 * line_dbl, f <- f * line_add over the bits of |x|, x = -0xd201000000010000; f <- conj(f) for x < 0
//   final    f^((q^12 - 1) / r) by plain square-and-multiply with the exponent computed once from q and r (no Frobenius
//            tables to get wrong; ~30 ms per pairing product, which a verifier pays once or twice per proof)
// Affine coordinates on the twist (one Fq inversion per step): a Miller loop is 68 steps.
#pragma once
#include <vector>
#include "host_math.hpp"

namespace zp {
namespace host {

struct Fq2 {
    Fq c0, c1;
    static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
    Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
    Fq2 neg() const { return {c0.neg(), c1.neg()}; }
    Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    Fq2 operator*(const Fq2& o) const {  // (a + bu)(c + du) = (ac - bd) + ((a + b)(c + d) - ac - bd) u
        Fq ac = c0 * o.c0, bd = c1 * o.c1;
        return {ac - bd, (c0 + c1) * (o.c0 + o.c1) - ac - bd};
    }
    Fq2 sqr() const { return *this * *this; }
    Fq2 mul_fq(const Fq& s) const { return {c0 * s, c1 * s}; }
    Fq2 mul_xi() const { return {c0 - c1, c0 + c1}; }  // (a + bu)(1 + u)
    Fq2 inverse() const {                              // conj / norm
        Fq n = (c0.sqr() + c1.sqr()).inverse();
        return {c0 * n, (c1 * n).neg()};
    }
};

struct Fq6 {
    Fq2 c0, c1, c2;
    static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
    static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
    bool operator==(const Fq6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
    Fq6 operator+(const Fq6& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
    Fq6 operator-(const Fq6& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
    Fq6 mul_v() const { return {c2.mul_xi(), c0, c1}; }  // v * (a + bv + cv^2) = c xi + a v + b v^2
    Fq6 operator*(const Fq6& o) const {                  // schoolbook with v^3 = xi (9 Fq2 products; clarity over speed)
        Fq2 t0 = c0 * o.c0 + (c1 * o.c2 + c2 * o.c1).mul_xi();
        Fq2 t1 = c0 * o.c1 + c1 * o.c0 + (c2 * o.c2).mul_xi();
        Fq2 t2 = c0 * o.c2 + c1 * o.c1 + c2 * o.c0;
        return {t0, t1, t2};
    }
};

struct Fq12 {
    Fq6 c0, c1;
    static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
    bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
    bool is_one() const { return *this == one(); }
    Fq12 operator*(const Fq12& o) const {  // (a + bw)(c + dw) = (ac + bd v) + ((a + b)(c + d) - ac - bd) w
        Fq6 ac = c0 * o.c0, bd = c1 * o.c1;
        return {ac + bd.mul_v(), (c0 + c1) * (o.c0 + o.c1) - ac - bd};
    }
    Fq12 sqr() const {  // complex squaring: 2 Fq6 products
        Fq6 t = c0 * c1;
        return {(c0 + c1) * (c0 + c1.mul_v()) - t - t.mul_v(), t + t};
    }
    Fq12 conj() const { return {c0, Fq6::zero() - c1}; }  // f^(q^6)
    // the sparse line value a0 + a1 v + a4 v w
    static Fq12 line(const Fq2& a0, const Fq2& a1, const Fq2& a4) {
        return {{a0, a1, Fq2::zero()}, {Fq2::zero(), a4, Fq2::zero()}};
    }
};

struct G2Affine {
    Fq2 x, y;
    bool inf;
};

// standard generator of G2, Montgomery limbs (same point as "Prize 1B/plonk-core/lib/blst/src/e2.c":23-45 BLS12_381_G2)
static inline G2Affine g2_generator() {
    static const uint64_t X0[6] = {0xf5f28fa202940a10ULL, 0xb3f5fb2687b4961aULL, 0xa1a893b53e2ae580ULL,
                                   0x9894999d1a3caee9ULL, 0x6f67b7631863366bULL, 0x058191924350bcd7ULL};
    static const uint64_t X1[6] = {0xa5a9c0759e23f606ULL, 0xaaa0c59dbccd60c3ULL, 0x3bb17e18e2867806ULL,
                                   0x1b1ab6cc8541b367ULL, 0xc2b6ed0ef2158547ULL, 0x11922a097360edf3ULL};
    static const uint64_t Y0[6] = {0x4c730af860494c4aULL, 0x597cfa1f5e369c5aULL, 0xe7e6856caa0a635aULL,
                                   0xbbefb5e96e0d495fULL, 0x07d3a975f0ef25a2ULL, 0x0083fd8e7e80dae5ULL};
    static const uint64_t Y1[6] = {0xadc0fc92df64b05dULL, 0x18aa270a2b1461dcULL, 0x86adac6a3be4eba0ULL,
                                   0x79495c4ec93da33aULL, 0xe7175850a43ccaedULL, 0x0b2bc2a163de1bf2ULL};
    G2Affine g;
    memcpy(g.x.c0.v, X0, 48);
    memcpy(g.x.c1.v, X1, 48);
    memcpy(g.y.c0.v, Y0, 48);
    memcpy(g.y.c1.v, Y1, 48);
    g.inf = false;
    return g;
}
static inline bool g2_on_curve(const G2Affine& p) {
    if (p.inf) return true;
    Fq2 b = Fq2{Fq::from_u64(4), Fq::zero()}.mul_xi();
    return p.y.sqr() == p.x.sqr() * p.x + b;
}
// affine chord / tangent addition on the twist (inputs finite)
static inline G2Affine g2_add(const G2Affine& a, const G2Affine& b) {
    if (a.inf) return b;
    if (b.inf) return a;
    Fq2 lam;
    if (a.x == b.x) {
        if (!(a.y == b.y) || a.y.is_zero()) return {Fq2::zero(), Fq2::zero(), true};
        Fq2 xx = a.x.sqr();
        lam = (xx.dbl() + xx) * a.y.dbl().inverse();
    } else {
        lam = (b.y - a.y) * (b.x - a.x).inverse();
    }
    Fq2 x3 = lam.sqr() - a.x - b.x;
    return {x3, lam * (a.x - x3) - a.y, false};
}
static inline G2Affine g2_mul(const G2Affine& p, const Fr& s) {  // scalar in Montgomery form
    uint64_t k[4];
    s.to_canonical(k);
    G2Affine acc{Fq2::zero(), Fq2::zero(), true};
    for (int i = 254; i >= 0; i--) {
        acc = g2_add(acc, acc);
        if ((k[i >> 6] >> (i & 63)) & 1) acc = g2_add(acc, p);
    }
    return acc;
}

static const uint64_t BLS_X_ABS = 0xd201000000010000ULL;  // |x|, x negative

// f_{|x|, Q}(P), conjugated for the sign of x.  P = (px, py) finite affine G1, Q finite affine G2.
static inline Fq12 miller_loop(const Fq& px, const Fq& py, const G2Affine& q) {
    Fq12 f = Fq12::one();
    Fq2 tx = q.x, ty = q.y;
    const Fq2 ypv{py, Fq::zero()};
    const Fq npx = px.neg();
    for (int i = 62; i >= 0; i--) {  // bit 63 of |x| is the leading one
        Fq2 xx = tx.sqr();
        Fq2 lam = (xx.dbl() + xx) * ty.dbl().inverse();
        f = f.sqr() * Fq12::line(lam * tx - ty, lam.mul_fq(npx), ypv);
        Fq2 x3 = lam.sqr() - tx.dbl();
        ty = lam * (tx - x3) - ty;
        tx = x3;
        if ((BLS_X_ABS >> i) & 1) {
            lam = (ty - q.y) * (tx - q.x).inverse();
            f = f * Fq12::line(lam * tx - ty, lam.mul_fq(npx), ypv);
            x3 = lam.sqr() - tx - q.x;
            ty = lam * (tx - x3) - ty;
            tx = x3;
        }
    }
    return f.conj();
}

// ---- (q^12 - 1) / r, computed once with schoolbook big integers
namespace big {
typedef std::vector<uint64_t> N;
static inline N mul(const N& a, const N& b) {
    N r(a.size() + b.size(), 0);
    for (size_t i = 0; i < a.size(); i++) {
        u128 c = 0;
        for (size_t j = 0; j < b.size(); j++) {
            c += (u128)a[i] * b[j] + r[i + j];
            r[i + j] = (uint64_t)c;
            c >>= 64;
        }
        r[i + b.size()] = (uint64_t)c;
    }
    return r;
}
static inline int cmp(const N& a, const N& b) {  // equal length
    for (size_t i = a.size(); i-- > 0;)
        if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return 0;
}
// floor(a / d), bit-serial; the remainder is returned through rem
static inline N div(const N& a, const N& d, N* rem) {
    N q(a.size(), 0), r(d.size() + 1, 0), dd(d);
    dd.push_back(0);
    for (size_t bit = a.size() * 64; bit-- > 0;) {
        for (size_t i = r.size(); i-- > 1;) r[i] = (r[i] << 1) | (r[i - 1] >> 63);
        r[0] = (r[0] << 1) | ((a[bit >> 6] >> (bit & 63)) & 1);
        if (cmp(r, dd) >= 0) {
            uint64_t br = 0;
            for (size_t i = 0; i < r.size(); i++) {
                u128 t = (u128)r[i] - dd[i] - br;
                r[i] = (uint64_t)t;
                br = (uint64_t)(t >> 64) & 1;
            }
            q[bit >> 6] |= (uint64_t)1 << (bit & 63);
        }
    }
    if (rem) *rem = r;
    return q;
}
}  // namespace big

static inline const big::N& final_exponent() {
    static const big::N e = [] {
        big::N q(Params<6>::p(), Params<6>::p() + 6), r(Params<4>::p(), Params<4>::p() + 4);
        big::N p = q;
        for (int i = 1; i < 12; i++) p = big::mul(p, q);
        p[0] -= 1;  // q^12 is odd
        big::N rem;
        big::N res = big::div(p, r, &rem);
        for (uint64_t w : rem)
            if (w) abort();  // r | q^12 - 1
        while (!res.empty() && res.back() == 0) res.pop_back();
        return res;
    }();
    return e;
}
static inline Fq12 final_exponentiation(const Fq12& f) {
    const big::N& e = final_exponent();
    Fq12 r = Fq12::one();
    bool started = false;
    for (size_t bit = e.size() * 64; bit-- > 0;) {
        if (started) r = r.sqr();
        if ((e[bit >> 6] >> (bit & 63)) & 1) {
            r = started ? r * f : f;
            started = true;
        }
    }
    return r;
}

// prod_i e(P_i, Q_i) == 1 ?   (pairs with P_i or Q_i at infinity contribute 1)
struct PairingInput {
    Fq px, py;
    bool p_inf;
    G2Affine q;
};
static inline Fq12 pairing_product(const std::vector<PairingInput>& in) {
    Fq12 f = Fq12::one();
    for (auto& t : in)
        if (!t.p_inf && !t.q.inf) f = f * miller_loop(t.px, t.py, t.q);
    return final_exponentiation(f);
}

}  // namespace host
}  // namespace zp
