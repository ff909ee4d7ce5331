// Host-side scalar arithmetic of the prover driver: the handful of Fr operations per proof that depend
// on Fiat-Shamir challenges (z^N, L_1(z), products of evaluations …) and the Fq/G1 tail of each MSM
// (Horner combination of the per-window sums, conversion to affine).  The reference does the same tails
// on the CPU ("Prize 1B/plonk-core/lib/PLONK/utils/zkp/cpu/collect.h":378-488, "…/src/point.cu":29-47).
// 64-bit limbs + unsigned __int128; byte layout identical to the device types (Montgomery, LE).
#pragma once
#include <stdint.h>
#include <string.h>
#include "curve.cuh"

namespace zp {
namespace host {

typedef unsigned __int128 u128;

template <int N>
struct Params;
template <>
struct Params<4> {
    static const uint64_t* p() {
        static const uint64_t v[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
        return v;
    }
    static const uint64_t* one() {
        static const uint64_t v[4] = {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL};
        return v;
    }
    static const uint64_t* rr() {
        static const uint64_t v[4] = {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL};
        return v;
    }
    static constexpr uint64_t inv = 0xfffffffeffffffffULL;
};
template <>
struct Params<6> {
    static const uint64_t* p() {
        static const uint64_t v[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                      0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
        return v;
    }
    static const uint64_t* one() {
        static const uint64_t v[6] = {0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                                      0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL};
        return v;
    }
    static const uint64_t* rr() {
        static const uint64_t v[6] = {0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL,
                                      0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL};
        return v;
    }
    static constexpr uint64_t inv = 0x89f3fffcfffcfffdULL;
};

template <int N>
struct F {
    uint64_t v[N];
    static F zero() { F r; memset(r.v, 0, sizeof(r.v)); return r; }
    static F one() { F r; memcpy(r.v, Params<N>::one(), sizeof(r.v)); return r; }
    bool is_zero() const { uint64_t a = 0; for (int i = 0; i < N; i++) a |= v[i]; return a == 0; }
    bool operator==(const F& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
    bool operator!=(const F& o) const { return !(*this == o); }
    static bool geq_p(const uint64_t* a) {
        const uint64_t* p = Params<N>::p();
        for (int i = N - 1; i >= 0; i--) if (a[i] != p[i]) return a[i] > p[i];
        return true;
    }
    static void sub_p(uint64_t* a) {
        const uint64_t* p = Params<N>::p();
        uint64_t br = 0;
        for (int i = 0; i < N; i++) { u128 d = (u128)a[i] - p[i] - br; a[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    }
    F operator+(const F& o) const {
        F r; u128 c = 0;
        for (int i = 0; i < N; i++) { c += (u128)v[i] + o.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
        if (c || geq_p(r.v)) sub_p(r.v);
        return r;
    }
    F operator-(const F& o) const {
        F r; uint64_t br = 0;
        for (int i = 0; i < N; i++) { u128 d = (u128)v[i] - o.v[i] - br; r.v[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
        if (br) { const uint64_t* p = Params<N>::p(); u128 c = 0; for (int i = 0; i < N; i++) { c += (u128)r.v[i] + p[i]; r.v[i] = (uint64_t)c; c >>= 64; } }
        return r;
    }
    F neg() const { return zero() - *this; }
    F dbl() const { return *this + *this; }
    F operator*(const F& o) const {  // word-serial Montgomery (separate product and reduction sweeps)
        const uint64_t* p = Params<N>::p();
        uint64_t t[2 * N + 1];
        memset(t, 0, sizeof(t));
        for (int i = 0; i < N; i++) {
            u128 c = 0;
            for (int j = 0; j < N; j++) { c += (u128)v[j] * o.v[i] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
            t[i + N] = (uint64_t)c;
        }
        uint64_t top = 0;
        for (int i = 0; i < N; i++) {
            uint64_t m = t[i] * Params<N>::inv;
            u128 c = 0;
            for (int j = 0; j < N; j++) { c += (u128)m * p[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
            for (int j = i + N; j < 2 * N && c; j++) { c += t[j]; t[j] = (uint64_t)c; c >>= 64; }
            top += (uint64_t)c;
        }
        F r;
        memcpy(r.v, t + N, sizeof(r.v));
        if (top || geq_p(r.v)) sub_p(r.v);
        return r;
    }
    F sqr() const { return *this * *this; }
    F pow(const uint64_t* e, int n) const {
        F r = one();
        for (int i = n * 64 - 1; i >= 0; i--) { r = r.sqr(); if ((e[i / 64] >> (i % 64)) & 1) r = r * *this; }
        return r;
    }
    F pow_u64(uint64_t e) const { return pow(&e, 1); }
    F inverse() const {
        uint64_t e[N];
        memcpy(e, Params<N>::p(), sizeof(e));
        // p - 2 (p is odd; borrow handled generally)
        uint64_t br = 2;
        for (int i = 0; i < N && br; i++) { uint64_t o = e[i]; e[i] = o - br; br = o < br ? 1 : 0; }
        return pow(e, N);
    }
    static F from_u64(uint64_t x) {
        F a = zero(), rr;
        a.v[0] = x;
        memcpy(rr.v, Params<N>::rr(), sizeof(rr.v));
        return a * rr;
    }
    static F from_canonical(const uint64_t* c) {
        F a, rr;
        memcpy(a.v, c, sizeof(a.v));
        memcpy(rr.v, Params<N>::rr(), sizeof(rr.v));
        return a * rr;
    }
    void to_canonical(uint64_t* out) const {
        F o = zero();
        o.v[0] = 1;
        F r = *this * o;
        memcpy(out, r.v, sizeof(r.v));
    }
};
typedef F<4> Fr;
typedef F<6> Fq;

static inline Fr to_host(const fr_t& a) { Fr r; memcpy(r.v, a.l, 32); return r; }
static inline fr_t to_dev(const Fr& a) { fr_t r; memcpy(r.l, a.v, 32); return r; }
static inline Fq to_host(const fq_t& a) { Fq r; memcpy(r.v, a.l, 48); return r; }
static inline fq_t to_dev(const Fq& a) { fq_t r; memcpy(r.l, a.v, 48); return r; }

// XYZZ point on the host (same formulas as curve.cuh)
struct G1 {
    Fq X, Y, ZZ, ZZZ;
    static G1 infinity() { G1 r; r.X = r.Y = r.ZZ = r.ZZZ = Fq::zero(); return r; }
    bool is_inf() const { return ZZ.is_zero(); }
    static G1 from_dev(const xyzz_t& p) { G1 r; r.X = to_host(p.X); r.Y = to_host(p.Y); r.ZZ = to_host(p.ZZ); r.ZZZ = to_host(p.ZZZ); return r; }
    static G1 from_affine(const Fq& x, const Fq& y) { G1 r; r.X = x; r.Y = y; r.ZZ = r.ZZZ = Fq::one(); return r; }
    void dbl_inplace() {
        if (is_inf()) return;
        Fq U = Y.dbl(), V = U.sqr(), W = U * V, S = X * V, xx = X.sqr(), M = xx.dbl() + xx;
        Fq X3 = M.sqr() - S.dbl(), Y3 = M * (S - X3) - W * Y;
        X = X3; Y = Y3; ZZ = V * ZZ; ZZZ = W * ZZZ;
    }
    void add(const G1& o) {
        if (o.is_inf()) return;
        if (is_inf()) { *this = o; return; }
        Fq U1 = X * o.ZZ, U2 = o.X * ZZ, S1 = Y * o.ZZZ, S2 = o.Y * ZZZ, P = U2 - U1, R = S2 - S1;
        if (P.is_zero()) { if (R.is_zero()) dbl_inplace(); else *this = infinity(); return; }
        Fq PP = P.sqr(), PPP = P * PP, Q = U1 * PP;
        Fq X3 = R.sqr() - PPP - Q.dbl(), Y3 = R * (Q - X3) - S1 * PPP;
        X = X3; Y = Y3; ZZ = ZZ * o.ZZ * PP; ZZZ = ZZZ * o.ZZZ * PPP;
    }
    // affine (x, y); infinity is encoded as (0, Mont(1)) like the reference ("…/src/point.cu":30-34)
    void to_affine(Fq& x, Fq& y, bool& inf) const {
        if (is_inf()) { x = Fq::zero(); y = Fq::one(); inf = true; return; }
        // 1/ZZZ gives both: 1/ZZ = ZZ^2 / ZZZ^2 ... use two inversions folded into one: i = 1/(ZZ*ZZZ)
        Fq i = (ZZ * ZZZ).inverse();
        Fq zz_inv = i * ZZZ, zzz_inv = i * ZZ;
        x = X * zz_inv; y = Y * zzz_inv; inf = false;
    }
};

}  // namespace host
}  // namespace zp
