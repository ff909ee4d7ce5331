// BLS12-381 Fr / Fq Montgomery arithmetic on 32-bit limbs (8 limbs Fr, 12 limbs Fq).
//
// Replaces the reference's sppark `mont_t` ("Prize 1B/plonk-core/lib/PLONK/utils/mont/cuda/ff/
// mont_t.cuh":31-1142, constants "…/ff/bls12-381.hpp":7-93).  Same external contract: little-endian
// limbs, Montgomery form with R = 2^(32 N), every result canonical (< p), so device buffers are
// bit-identical to the ark-ff `Fp256/Fp384` values that cross the FFI.
//
// Multiplication is a CIOS loop whose row operation  acc[0..N] += x[0..N) * y  is written as two carry
// chains (even limbs, then odd limbs) of mad.lo.cc / madc.hi.cc so that every 32x32->64 product is one
// IMAD.WIDE with carry-in/out: 2N multiply-adds per row, 2 rows per limb of b  =>  4N^2 + N integer
// multiply-adds per product (264 for Fr, 588 for Fq — the figure SURVEY §8d uses for the MSM roofline).
// Because p < 2^(32N-1) the accumulator never needs more than N+1 limbs.
#pragma once
#include "ptx_ops.cuh"

namespace zp {

struct FrParams {
    static constexpr int N = 8;
    static constexpr uint32_t M0 = 0xffffffffu;  // -p^-1 mod 2^32
    ZP_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t v[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                   0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return v[i];
    }
    ZP_HD static constexpr uint32_t one(int i) {  // R mod p
        constexpr uint32_t v[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                                   0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return v[i];
    }
    ZP_HD static constexpr uint32_t rr(int i) {  // R^2 mod p
        constexpr uint32_t v[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                                   0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return v[i];
    }
};

struct FqParams {
    static constexpr int N = 12;
    static constexpr uint32_t M0 = 0xfffcfffdu;
    ZP_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t v[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                    0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return v[i];
    }
    ZP_HD static constexpr uint32_t one(int i) {
        constexpr uint32_t v[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                    0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return v[i];
    }
    ZP_HD static constexpr uint32_t rr(int i) {
        constexpr uint32_t v[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                    0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return v[i];
    }
};

template <class P>
struct Mont {
    static constexpr int N = P::N;
    uint32_t l[N];

    ZP_HD static Mont zero() {
        Mont r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    ZP_HD static Mont one() {
        Mont r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::one(i);
        return r;
    }
    ZP_HD static Mont rr() {
        Mont r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::rr(i);
        return r;
    }
    ZP_HD bool is_zero() const {
        uint32_t a = 0;
#pragma unroll
        for (int i = 0; i < N; i++) a |= l[i];
        return a == 0;
    }
    ZP_HD bool operator==(const Mont& o) const {
        uint32_t a = 0;
#pragma unroll
        for (int i = 0; i < N; i++) a |= l[i] ^ o.l[i];
        return a == 0;
    }
    ZP_HD bool operator!=(const Mont& o) const { return !(*this == o); }

    // r = (t >= p) ? t - p : t      (t < 2p)
    ZP_HD static void final_sub(uint32_t* t) {
        uint32_t s[N];
        s[0] = sub_cc(t[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < N; i++) s[i] = subc_cc(t[i], P::mod(i));
        uint32_t borrow = subc(0u, 0u);  // 0xffffffff if t < p
#pragma unroll
        for (int i = 0; i < N; i++) t[i] = borrow ? t[i] : s[i];
    }

    ZP_HD Mont operator+(const Mont& o) const {
        Mont r;
        r.l[0] = add_cc(l[0], o.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(l[i], o.l[i]);
        r.l[N - 1] = addc(l[N - 1], o.l[N - 1]);  // 2p < 2^(32N): no carry out
        final_sub(r.l);
        return r;
    }
    ZP_HD Mont operator-(const Mont& o) const {
        Mont r;
        r.l[0] = sub_cc(l[0], o.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.l[i] = subc_cc(l[i], o.l[i]);
        uint32_t borrow = subc(0u, 0u);
        // add p back under mask
        r.l[0] = add_cc(r.l[0], borrow & P::mod(0));
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(r.l[i], borrow & P::mod(i));
        r.l[N - 1] = addc(r.l[N - 1], borrow & P::mod(N - 1));
        return r;
    }
    ZP_HD Mont neg() const {
        Mont z = zero();
        return z - *this;
    }
    ZP_HD Mont dbl() const { return *this + *this; }

    // acc[0..N] += x[0..N) * y   — two carry chains, see header comment
    template <class X>
    ZP_HD static void row_mad(uint32_t* acc, X x, uint32_t y) {
        acc[0] = mad_lo_cc(x(0), y, acc[0]);
        acc[1] = madc_hi_cc(x(0), y, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            acc[j] = madc_lo_cc(x(j), y, acc[j]);
            acc[j + 1] = madc_hi_cc(x(j), y, acc[j + 1]);
        }
        acc[N] = addc(acc[N], 0u);
        acc[1] = mad_lo_cc(x(1), y, acc[1]);
        acc[2] = madc_hi_cc(x(1), y, acc[2]);
#pragma unroll
        for (int j = 3; j < N - 1; j += 2) {
            acc[j] = madc_lo_cc(x(j), y, acc[j]);
            acc[j + 1] = madc_hi_cc(x(j), y, acc[j + 1]);
        }
        acc[N - 1] = madc_lo_cc(x(N - 1), y, acc[N - 1]);
        acc[N] = madc_hi(x(N - 1), y, acc[N]);
    }

    struct LimbsOf {
        const uint32_t* p;
        ZP_HD uint32_t operator()(int i) const { return p[i]; }
    };
    struct ModLimbs {
        ZP_HD uint32_t operator()(int i) const { return P::mod(i); }
    };

    // ---- even/odd accumulator product -------------------------------------------------------
    // value = sum_k ev[k] 2^(32k) + sum_k od[k] 2^(32(k+1)).  Both arrays are only ever touched as
    // aligned limb pairs (ev[2t],ev[2t+1]) / (od[2t],od[2t+1]), so every mad.lo.cc + madc.hi.cc pair maps
    // to one IMAD.WIDE.U32.X on an even-aligned register pair.  Dividing by 2^32 after each reduction
    // step swaps the roles of the two arrays instead of moving limbs.
    //
    // ev/od += x * y, x given by functor (even limbs feed ev, odd limbs feed od).
    // `lone` (a limb of weight 2^0 left over from the previous shift) is folded into ev[0] first and
    // its carry enters the od chain (weight 2^32), exactly where it belongs.
    template <class X>
    ZP_HD static void eo_row(uint32_t* ev, uint32_t* od, X x, uint32_t y, bool with_lone, uint32_t lone) {
        if (with_lone) {
            ev[0] = add_cc(ev[0], lone);
            od[0] = madc_lo_cc(x(1), y, od[0]);
        } else {
            od[0] = mad_lo_cc(x(1), y, od[0]);
        }
        od[1] = madc_hi_cc(x(1), y, od[1]);
#pragma unroll
        for (int j = 3; j < N; j += 2) {
            od[j - 1] = madc_lo_cc(x(j), y, od[j - 1]);
            od[j] = madc_hi_cc(x(j), y, od[j]);
        }
        // no carry out of od[N-1]: od <= value / 2^32 < 2^(32N)
        ev[0] = mad_lo_cc(x(0), y, ev[0]);
        ev[1] = madc_hi_cc(x(0), y, ev[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            ev[j] = madc_lo_cc(x(j), y, ev[j]);
            ev[j + 1] = madc_hi_cc(x(j), y, ev[j + 1]);
        }
        od[N - 1] = addc(od[N - 1], 0u);  // carry out of ev (weight 2^(32N)) lives in od[N-1]
    }

    ZP_HD Mont operator*(const Mont& o) const {
        uint32_t A[N], B[N];  // the two accumulators; which one is "even" alternates per iteration
#pragma unroll
        for (int i = 0; i < N; i++) A[i] = B[i] = 0;
        uint32_t lone = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint32_t* ev = (i & 1) ? B : A;
            uint32_t* od = (i & 1) ? A : B;
            eo_row(ev, od, LimbsOf{l}, o.l[i], i != 0, lone);
            uint32_t m = ev[0] * P::M0;
            eo_row(ev, od, ModLimbs{}, m, false, 0u);
            // divide by 2^32: ev[0] == 0 now; ev[1] becomes the lone limb of weight 1;
            // ev[2..] slides down to become the next "odd" array, od becomes the next "even" array.
            lone = ev[1];
#pragma unroll
            for (int j = 0; j < N - 2; j++) ev[j] = ev[j + 2];
            ev[N - 2] = 0;
            ev[N - 1] = 0;
        }
        // after N iterations (N even) roles are back: A is "even", B is "odd"
        uint32_t acc[N];
        acc[0] = add_cc(A[0], lone);
#pragma unroll
        for (int k = 1; k < N; k++) acc[k] = addc_cc(A[k], B[k - 1]);
        final_sub(acc);
        Mont r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = acc[i];
        return r;
    }
    // (A dedicated squaring — off-diagonal products once, doubled, separate reduction sweep: 222 instead of 288 wide
    // products for Fq — was measured SLOWER on B200: 24.7 vs 30.7 G ops/s; the long carry ripples and shifts cost more
    // issue slots than the saved IMADs.  Squaring therefore goes through the multiplier.)
    ZP_HD Mont sqr() const { return *this * *this; }
    ZP_HD Mont pow5() const {
        Mont s = sqr();
        return s.sqr() * *this;
    }

    // Montgomery -> canonical: multiply by 1
    ZP_HD Mont from_mont() const {
        uint32_t acc[N + 1];
#pragma unroll
        for (int i = 0; i < N; i++) acc[i] = l[i];
        acc[N] = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint32_t m = acc[0] * P::M0;
            row_mad(acc, ModLimbs{}, m);
#pragma unroll
            for (int j = 0; j < N; j++) acc[j] = acc[j + 1];
            acc[N] = 0;
        }
        final_sub(acc);
        Mont r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = acc[i];
        return r;
    }
    ZP_HD Mont to_mont() const { return *this * rr(); }

    // generic exponentiation by a little-endian limb array (not constant time; public exponents only)
    ZP_HD Mont pow(const uint32_t* e, int nlimbs) const {
        Mont r = one();
        for (int i = nlimbs * 32 - 1; i >= 0; i--) {
            r = r.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) r = r * *this;
        }
        return r;
    }
    ZP_HD Mont pow_u64(uint64_t e) const {
        uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
        return pow(w, 2);
    }
    // Fermat inverse a^(p-2); inverse(0) = 0.
    ZP_HD Mont inverse() const {
        uint32_t e[N];
        e[0] = sub_cc(P::mod(0), 2u);
#pragma unroll
        for (int i = 1; i < N; i++) e[i] = subc_cc(P::mod(i), 0u);
        return pow(e, N);
    }
    ZP_HD static Mont from_u32(uint32_t x) {
        Mont r = zero();
        r.l[0] = x;
        return r.to_mont();
    }
    // canonical-integer comparison a > b (inputs canonical, i.e. already from_mont'ed)
    ZP_HD static bool gt_canonical(const Mont& a, const Mont& b) {
        for (int i = N - 1; i >= 0; i--) {
            if (a.l[i] != b.l[i]) return a.l[i] > b.l[i];
        }
        return false;
    }
};

typedef Mont<FrParams> fr_t;
typedef Mont<FqParams> fq_t;

}  // namespace zp
