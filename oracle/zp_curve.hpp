// ORACLE — TEST INFRASTRUCTURE ONLY (see zp_field.hpp header).
//
// BLS12-381 G1 (y^2 = x^3 + 4) on the CPU: Jacobian arithmetic, scalar multiplication, a plain
// bucket-method MSM.  Restates what the reference obtains from ark-ec 0.3.0
// (`VariableBaseMSM::multi_scalar_mul`, `GroupAffine`; call sites
// "Prize 1B/plonk-core/src/proof_system/prover.rs":221,297,320-325,369,395,465) and what PNP does in
// "Prize 1B/plonk-core/lib/PLONK/src/point.cu":29-258 (to_affine: infinity -> (0, Mont(1))).
// Only the affine result of any group computation is observable, so any exact algorithm is bit-exact.
#pragma once
#include "zp_field.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace zpo {

struct G1Affine {
    Fq x, y;
    bool inf;
    static G1Affine infinity() {
        G1Affine a;
        a.x = Fq::zero();
        a.y = Fq::one();  // PNP/FFI encoding of infinity: (0, Mont(1)), point.cu:30-34
        a.inf = true;
        return a;
    }
    bool operator==(const G1Affine& o) const {
        if (inf || o.inf) return inf == o.inf;
        return x == o.x && y == o.y;
    }
};

struct G1 {  // Jacobian, Z == 0 is infinity
    Fq X, Y, Z;
    static G1 infinity() {
        G1 r;
        r.X = Fq::zero();
        r.Y = Fq::one();
        r.Z = Fq::zero();
        return r;
    }
    static G1 from_affine(const G1Affine& a) {
        if (a.inf) return infinity();
        G1 r;
        r.X = a.x;
        r.Y = a.y;
        r.Z = Fq::one();
        return r;
    }
    bool is_inf() const { return Z.is_zero(); }

    G1 dbl() const {
        if (is_inf()) return *this;
        // dbl-2009-l (a = 0)
        Fq A = X.square(), B = Y.square(), C = B.square();
        Fq D = ((X + B).square() - A - C).dbl();
        Fq E = A.dbl() + A;
        Fq F = E.square();
        G1 r;
        r.X = F - D.dbl();
        r.Y = E * (D - r.X) - C.dbl().dbl().dbl();
        r.Z = (Y * Z).dbl();
        return r;
    }
    G1 add(const G1& o) const {
        if (is_inf()) return o;
        if (o.is_inf()) return *this;
        Fq Z1Z1 = Z.square(), Z2Z2 = o.Z.square();
        Fq U1 = X * Z2Z2, U2 = o.X * Z1Z1;
        Fq S1 = Y * o.Z * Z2Z2, S2 = o.Y * Z * Z1Z1;
        if (U1 == U2) {
            if (S1 == S2) return dbl();
            return infinity();
        }
        Fq H = U2 - U1, R = S2 - S1;
        Fq HH = H.square(), HHH = H * HH, V = U1 * HH;
        G1 r;
        r.X = R.square() - HHH - V.dbl();
        r.Y = R * (V - r.X) - S1 * HHH;
        r.Z = Z * o.Z * H;
        return r;
    }
    G1 add_affine(const G1Affine& a) const { return add(from_affine(a)); }
    G1 neg() const {
        G1 r = *this;
        r.Y = -r.Y;
        return r;
    }
    // scalar given as canonical 256-bit little-endian limbs
    G1 mul_canonical(const uint64_t* k) const {
        G1 r = infinity();
        for (int i = 255; i >= 0; i--) {
            r = r.dbl();
            if ((k[i / 64] >> (i % 64)) & 1) r = r.add(*this);
        }
        return r;
    }
    G1 mul(const Fr& s) const {
        uint64_t k[4];
        s.to_canonical(k);
        return mul_canonical(k);
    }
    G1Affine to_affine() const {
        if (is_inf()) return G1Affine::infinity();
        Fq zi = Z.inverse();
        Fq zi2 = zi.square();
        G1Affine a;
        a.x = X * zi2;
        a.y = Y * zi2 * zi;
        a.inf = false;
        return a;
    }
};

static inline bool g1_on_curve(const G1Affine& a) {
    if (a.inf) return true;
    return a.y.square() == a.x.square() * a.x + Fq::from_u64(4);
}

// Standard BLS12-381 G1 generator (ark-bls12-381 G1_GENERATOR_X/Y; same point as blst's BLS12_381_G1).
static inline G1Affine g1_generator() {
    G1Affine g;
    g.x = Fq::from_hex(
        "17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb");
    g.y = Fq::from_hex(
        "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1");
    g.inf = false;
    return g;
}

// Batch Jacobian -> affine (Montgomery's trick).
static inline void g1_batch_to_affine(const std::vector<G1>& in, std::vector<G1Affine>& out) {
    size_t n = in.size();
    out.resize(n);
    std::vector<Fq> prefix(n);
    Fq acc = Fq::one();
    for (size_t i = 0; i < n; i++) {
        prefix[i] = acc;
        if (!in[i].is_inf()) acc = acc * in[i].Z;
    }
    Fq inv = acc.inverse();
    for (size_t i = n; i-- > 0;) {
        if (in[i].is_inf()) {
            out[i] = G1Affine::infinity();
            continue;
        }
        Fq zi = inv * prefix[i];
        inv = inv * in[i].Z;
        Fq zi2 = zi.square();
        out[i].x = in[i].X * zi2;
        out[i].y = in[i].Y * zi2 * zi;
        out[i].inf = false;
    }
}

static inline bool g1_msm_uses_blst() { return blst_api().on; }

// Bucket-method MSM over canonical scalars (window c, unsigned digits), OpenMP over windows.
// Restates `VariableBaseMSM::multi_scalar_mul` (ark-ec 0.3.0): exact group sum.
static inline G1 g1_msm(const G1Affine* pts, const Fr* scalars, size_t n) {
    if (n == 0) return G1::infinity();
    if (blst_api().on && n >= 32) {
        // point-range slices, one blst Pippenger per host thread, partial sums added in slice order
        int threads = 1;
#ifdef _OPENMP
        threads = omp_get_max_threads();
#endif
        size_t per = (n + threads - 1) / threads;
        if (per < 16) per = 16;
        int slices = (int)((n + per - 1) / per);
        std::vector<G1> part(slices, G1::infinity());
        bool any_inf = false;
        for (size_t i = 0; i < n && !any_inf; i++) any_inf = pts[i].inf;
        if (!any_inf) {
#pragma omp parallel for schedule(static, 1)
            for (int t = 0; t < slices; t++) {
                size_t lo = (size_t)t * per, hi = std::min(n, lo + per), m = hi - lo;
                std::vector<uint64_t> xy(12 * m), sc(4 * m);
                for (size_t i = 0; i < m; i++) {
                    memcpy(&xy[12 * i], pts[lo + i].x.v, 48);
                    memcpy(&xy[12 * i + 6], pts[lo + i].y.v, 48);
                    scalars[lo + i].to_canonical(&sc[4 * i]);
                }
                std::vector<uint64_t> scratch(blst_api().scratch_sizeof(m) / 8 + 1);
                const void* pp[2] = {xy.data(), nullptr};
                const void* sp[2] = {sc.data(), nullptr};
                G1 r;
                blst_api().pippenger(&r, pp, m, sp, 255, scratch.data());
                part[t] = r;
            }
            G1 total = G1::infinity();
            for (int t = 0; t < slices; t++) total = total.add(part[t]);
            return total;
        }
    }
    int c = 3;
    while ((1ull << (c + 1)) * 4 < n && c < 15) c++;
    int nwin = (255 + c - 1) / c;
    std::vector<uint64_t> canon(4 * n);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) scalars[i].to_canonical(&canon[4 * i]);
    std::vector<G1> wsum(nwin);
#pragma omp parallel for schedule(dynamic, 1)
    for (int w = 0; w < nwin; w++) {
        std::vector<G1> buckets((size_t)1 << c, G1::infinity());
        int bit0 = w * c;
        for (size_t i = 0; i < n; i++) {
            const uint64_t* k = &canon[4 * i];
            uint64_t d = 0;
            for (int b = 0; b < c; b++) {
                int bit = bit0 + b;
                if (bit < 256) d |= ((k[bit / 64] >> (bit % 64)) & 1) << b;
            }
            if (d && !pts[i].inf) buckets[d] = buckets[d].add_affine(pts[i]);
        }
        G1 run = G1::infinity(), sum = G1::infinity();
        for (size_t d = ((size_t)1 << c) - 1; d >= 1; d--) {
            run = run.add(buckets[d]);
            sum = sum.add(run);
        }
        wsum[w] = sum;
    }
    G1 r = G1::infinity();
    for (int w = nwin - 1; w >= 0; w--) {
        for (int b = 0; b < c; b++) r = r.dbl();
        r = r.add(wsum[w]);
    }
    return r;
}

}  // namespace zpo
