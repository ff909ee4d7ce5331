// ORACLE — TEST INFRASTRUCTURE ONLY (see zp_field.hpp header).
//
// Radix-2 evaluation domains over Fr: restates ark-poly 0.3.0 `GeneralEvaluationDomain`
// {fft, ifft, coset_fft, coset_ifft, elements, evaluate_vanishing_polynomial} as used at
// "Prize 1B/plonk-core/src/proof_system/prover.rs":177-211, "…/quotient_poly.rs":64-148,204 and
// "…/permutation/mod.rs":669-674,751.  In-tree corroboration of the semantics (natural order in/out,
// 1/N inside ifft, coset shift by g = 7): "…/lib/PLONK/utils/function.cu":249-273,
// "…/lib/PLONK/utils/zkp/cuda/zksnark_ntt/ntt_kernel/kernels.cuh":116-140.
#pragma once
#include "zp_field.hpp"
#include <algorithm>

namespace zpo {

struct Domain {
    int logn;
    size_t n;
    Fr omega, omega_inv, n_inv;
    std::vector<Fr> tw, tw_inv;  // omega^i, i < n/2

    explicit Domain(int logn_) : logn(logn_), n((size_t)1 << logn_) {
        ensure_init();
        omega = fr_root_of_unity(logn);
        omega_inv = omega.inverse();
        n_inv = Fr::from_u64(n).inverse();
        tw.resize(std::max<size_t>(n / 2, 1));
        tw_inv.resize(tw.size());
        tw[0] = tw_inv[0] = Fr::one();
        // chunked so the table can be built in parallel
        const size_t CH = 1024;
        size_t half = tw.size();
        for (size_t i = 1; i < std::min(CH, half); i++) {
            tw[i] = tw[i - 1] * omega;
            tw_inv[i] = tw_inv[i - 1] * omega_inv;
        }
        if (half > CH) {
            Fr step = tw[CH - 1] * omega, step_inv = tw_inv[CH - 1] * omega_inv;
            std::vector<Fr> base(half / CH), base_inv(half / CH);
            base[0] = base_inv[0] = Fr::one();
            for (size_t k = 1; k < half / CH; k++) {
                base[k] = base[k - 1] * step;
                base_inv[k] = base_inv[k - 1] * step_inv;
            }
#pragma omp parallel for schedule(static)
            for (long k = 1; k < (long)(half / CH); k++)
                for (size_t i = 0; i < CH; i++) {
                    tw[k * CH + i] = base[k] * tw[i];
                    tw_inv[k * CH + i] = base_inv[k] * tw_inv[i];
                }
        }
    }

    Fr element(size_t i) const {
        i &= n - 1;
        if (i < n / 2 || n == 1) return tw[i];
        return -tw[i - n / 2];
    }

    void transform(std::vector<Fr>& a, const std::vector<Fr>& t) const {
        assert(a.size() == n);
        // bit reversal
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (long ii = 0; ii < (long)n; ii++) {
            size_t i = (size_t)ii;
            size_t j = 0;
            for (int b = 0; b < logn; b++) j |= ((i >> b) & 1) << (logn - 1 - b);
            if (i < j) std::swap(a[i], a[j]);
        }
        // the first LB stages touch blocks of 2^LB consecutive elements only: run them block by block (cache resident,
        // one thread per block), the remaining stages sweep the whole array
        const int LB = std::min(logn, 13);
        const size_t BS = (size_t)1 << LB;
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (long blk0 = 0; blk0 < (long)(n / BS); blk0++) {
            Fr* x = a.data() + (size_t)blk0 * BS;
            for (int s = 1; s <= LB; s++) {
                size_t m = (size_t)1 << s, half = m >> 1, stride = n / m;
                for (size_t b = 0; b < BS; b += m)
                    for (size_t j = 0; j < half; j++) {
                        Fr u = x[b + j], v = x[b + j + half] * t[j * stride];
                        x[b + j] = u + v;
                        x[b + j + half] = u - v;
                    }
            }
        }
        for (int s = LB + 1; s <= logn; s++) {
            size_t m = (size_t)1 << s, half = m >> 1, stride = n / m;
#pragma omp parallel for schedule(static) if (n >= 4096)
            for (long k = 0; k < (long)(n / 2); k++) {
                size_t blk = (size_t)k / half, j = (size_t)k % half;
                size_t i0 = blk * m + j, i1 = i0 + half;
                Fr u = a[i0], v = a[i1] * t[j * stride];
                a[i0] = u + v;
                a[i1] = u - v;
            }
        }
    }
    // evals[i] = p(omega^i); input zero-padded to n
    std::vector<Fr> fft(const std::vector<Fr>& coeffs) const {
        std::vector<Fr> a(coeffs);
        assert(a.size() <= n);
        a.resize(n, Fr::zero());
        transform(a, tw);
        return a;
    }
    std::vector<Fr> ifft(const std::vector<Fr>& evals) const {
        std::vector<Fr> a(evals);
        assert(a.size() <= n);
        a.resize(n, Fr::zero());
        transform(a, tw_inv);
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (long i = 0; i < (long)n; i++) a[i] = a[i] * n_inv;
        return a;
    }
    static void distribute_powers(std::vector<Fr>& a, const Fr& g) {
        size_t n = a.size();
        const size_t CH = 4096;
        size_t nch = (n + CH - 1) / CH;
        std::vector<Fr> base(nch);
        Fr gch = g.pow_u64(CH);
        base[0] = Fr::one();
        for (size_t k = 1; k < nch; k++) base[k] = base[k - 1] * gch;
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (long k = 0; k < (long)nch; k++) {
            Fr p = base[k];
            for (size_t i = k * CH; i < std::min(n, (size_t)(k + 1) * CH); i++) {
                a[i] = a[i] * p;
                p = p * g;
            }
        }
    }
    std::vector<Fr> coset_fft(const std::vector<Fr>& coeffs) const {
        std::vector<Fr> a(coeffs);
        assert(a.size() <= n);
        a.resize(n, Fr::zero());
        distribute_powers(a, fr_generator());
        transform(a, tw);
        return a;
    }
    std::vector<Fr> coset_ifft(const std::vector<Fr>& evals) const {
        std::vector<Fr> a = ifft(evals);
        distribute_powers(a, fr_generator().inverse());
        return a;
    }
    Fr evaluate_vanishing(const Fr& z) const { return z.pow_u64(n) - Fr::one(); }
};

// Horner evaluation (DensePolynomial::evaluate)
static inline Fr poly_eval(const std::vector<Fr>& c, const Fr& z) {
    size_t n = c.size();
    if (n < 8192) {
        Fr r = Fr::zero();
        for (size_t i = n; i-- > 0;) r = r * z + c[i];
        return r;
    }
    const size_t CH = 4096;
    size_t nch = (n + CH - 1) / CH;
    std::vector<Fr> part(nch);
#pragma omp parallel for schedule(static)
    for (long k = 0; k < (long)nch; k++) {
        Fr r = Fr::zero();
        size_t hi = std::min(n, (size_t)(k + 1) * CH);
        for (size_t i = hi; i-- > (size_t)k * CH;) r = r * z + c[i];
        part[k] = r;
    }
    Fr zch = z.pow_u64(CH), r = Fr::zero();
    for (size_t k = nch; k-- > 0;) r = r * zch + part[k];
    return r;
}

// floor(p(X) / (X - z)), remainder dropped  (ark-poly-commit kzg10 `compute_witness_polynomial`;
// PNP twin "…/lib/PLONK/src/KZG/kzg10.cu":87-99)
static inline std::vector<Fr> poly_div_linear(const std::vector<Fr>& p, const Fr& z) {
    size_t n = p.size();
    if (n <= 1) return std::vector<Fr>();
    std::vector<Fr> q(n - 1);
    Fr carry = Fr::zero();
    for (size_t i = n - 1; i >= 1; i--) {
        carry = p[i] + carry * z;
        q[i - 1] = carry;
    }
    return q;
}

static inline void batch_inverse(std::vector<Fr>& v) {
    size_t n = v.size();
    std::vector<Fr> pre(n);
    Fr acc = Fr::one();
    for (size_t i = 0; i < n; i++) {
        pre[i] = acc;
        acc = acc * v[i];
    }
    Fr inv = acc.inverse();
    for (size_t i = n; i-- > 0;) {
        Fr t = inv * pre[i];
        inv = inv * v[i];
        v[i] = t;
    }
}

}  // namespace zpo
