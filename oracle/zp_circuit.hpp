// ORACLE — TEST INFRASTRUCTURE ONLY (see zp_field.hpp header).
//
// Synthetic circuit front end + preprocessing, i.e. everything ABOVE the prover hot path that the
// reference does in Rust and that we only need in order to fabricate inputs in the exact FFI layout:
//   * a miniature StandardComposer (variables, 4 wires, 15 selector columns, copy-constraint map):
//     "Prize 1B/plonk-core/src/constraint_system/composer.rs":241-367,604-679,
//     "…/constraint_system/hash.rs":20-127 (degree-5 Poseidon gates);
//   * a Poseidon-shaped Merkle-tree circuit with the reference's gate census (1 zero gate + 3 blinding
//     rows + 193 gates per hash + 1 public-input gate) — synthetic MDS/round constants;
//   * sigma construction: "…/permutation/mod.rs":101-215;
//   * prover-key construction: "…/proof_system/preprocess.rs":64-99,162-295,498-520 and
//     "…/lookup/preprocess.rs":41-67, "…/lookup/multiset.rs":70-79;
//   * KZG10 setup with a KNOWN trapdoor tau (powers_of_g[i] = tau^i * G) so that commitments and
//     openings can be checked without pairings (SURVEY §8c pin (1)).
#pragma once
#include "zp_poly.hpp"
#include "zp_curve.hpp"
#include <array>
#include <utility>

namespace zpo {

enum Sel {
    Q_M = 0, Q_L, Q_R, Q_O, Q_4, Q_C, Q_HL, Q_HR, Q_H4, Q_ARITH, Q_RANGE, Q_LOGIC, Q_FIXED, Q_VAR, Q_LOOKUP,
    SIG_L, SIG_R, SIG_O, SIG_4, NUM_PK_POLYS
};
static const int NUM_SELECTORS = 15;

static inline Fr K_const(int wire) {  // permutation/constants.rs:12-22
    static const uint64_t k[4] = {1, 7, 13, 17};
    return Fr::from_u64(k[wire]);
}

struct Composer {
    std::vector<Fr> var_vals;
    std::vector<uint32_t> w[4];
    std::vector<Fr> q[NUM_SELECTORS];
    std::vector<std::pair<uint64_t, Fr>> pi;  // non-zero public inputs (pos, value)
    std::vector<std::array<Fr, 4>> table;      // lookup table rows
    uint32_t zero_var;

    size_t n() const { return w[0].size(); }
    uint32_t add_input(const Fr& v) {
        var_vals.push_back(v);
        return (uint32_t)var_vals.size() - 1;
    }
    void push_gate(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        w[0].push_back(a);
        w[1].push_back(b);
        w[2].push_back(c);
        w[3].push_back(d);
        for (int s = 0; s < NUM_SELECTORS; s++) q[s].push_back(Fr::zero());
    }
    Fr& sel(int s) { return q[s].back(); }

    // composer.rs:275-318
    void poly_gate(uint32_t a, uint32_t b, uint32_t c, Fr qm, Fr ql, Fr qr, Fr qo, Fr qc, const Fr* pival) {
        push_gate(a, b, c, zero_var);
        sel(Q_M) = qm;
        sel(Q_L) = ql;
        sel(Q_R) = qr;
        sel(Q_O) = qo;
        sel(Q_C) = qc;
        sel(Q_ARITH) = Fr::one();
        if (pival && !pival->is_zero()) pi.push_back({(uint64_t)n() - 1, *pival});
    }
    // composer.rs:241-246 + 604-679, blinding values drawn from the seeded test RNG
    void prelude(SplitMix64& rng) {
        zero_var = add_input(Fr::zero());
        poly_gate(zero_var, zero_var, zero_var, Fr::zero(), Fr::one(), Fr::zero(), Fr::zero(), Fr::zero(), nullptr);
        uint32_t r1 = zero_var, r2 = zero_var;
        for (int i = 0; i < 2; i++) {
            r1 = add_input(rng.next_fr());
            r2 = add_input(rng.next_fr());
            uint32_t r3 = add_input(rng.next_fr());
            uint32_t r4 = add_input(rng.next_fr());
            push_gate(r1, r2, r3, r4);
        }
        push_gate(r1, r2, zero_var, zero_var);
    }
    // division by -q_o: the Poseidon gadget always uses q_o = -1, so cache the last inverse
    Fr inv_last_in = Fr::zero(), inv_last_out = Fr::zero();
    Fr inv_cached(const Fr& x) {
        if (x == Fr::one()) return x;
        if (!(x == inv_last_in)) {
            inv_last_in = x;
            inv_last_out = x.inverse();
        }
        return inv_last_out;
    }
    // hash.rs:23-67
    uint32_t full_affine_transform_gate(const uint32_t v[3], const Fr s[5]) {
        Fr val = (s[0] * var_vals[v[0]].pow_u64(5) + s[1] * var_vals[v[1]].pow_u64(5) +
                  s[2] * var_vals[v[2]].pow_u64(5) + s[3]) *
                 inv_cached(-s[4]);
        uint32_t o = add_input(val);
        push_gate(v[0], v[1], o, v[2]);
        sel(Q_HL) = s[0];
        sel(Q_HR) = s[1];
        sel(Q_H4) = s[2];
        sel(Q_C) = s[3];
        sel(Q_O) = s[4];
        sel(Q_ARITH) = Fr::one();
        return o;
    }
    // hash.rs:76-120
    uint32_t partial_affine_transform_gate(const uint32_t v[3], const Fr s[5]) {
        Fr val = (s[0] * var_vals[v[0]].pow_u64(5) + s[1] * var_vals[v[1]] + s[2] * var_vals[v[2]] + s[3]) *
                 inv_cached(-s[4]);
        uint32_t o = add_input(val);
        push_gate(v[0], v[1], o, v[2]);
        sel(Q_HL) = s[0];
        sel(Q_R) = s[1];
        sel(Q_4) = s[2];
        sel(Q_C) = s[3];
        sel(Q_O) = s[4];
        sel(Q_ARITH) = Fr::one();
        return o;
    }
    void assert_equal(uint32_t a, uint32_t b) {  // composer.rs:355-367
        poly_gate(a, b, zero_var, Fr::zero(), Fr::one(), -Fr::one(), Fr::zero(), Fr::zero(), nullptr);
    }
    // plookup gate: q_lookup = 1, every arithmetic selector 0
    void lookup_gate(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        push_gate(a, b, c, d);
        sel(Q_LOOKUP) = Fr::one();
    }
};

// Poseidon-shaped permutation, width 3, R_F = 8, R_P = 55, alpha = 5, synthetic constants.
struct HashParams {
    Fr mds[3][3];
    Fr rc[63][3];
    Fr ark0[3];
    explicit HashParams(uint64_t seed) {
        SplitMix64 rng(seed);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) mds[i][j] = rng.next_fr();
        for (int r = 0; r < 63; r++)
            for (int j = 0; j < 3; j++) rc[r][j] = rng.next_fr();
        for (int j = 0; j < 3; j++) ark0[j] = rng.next_fr();
    }
};

// 193 gates: 3 addi + 63 rounds * 3 + 1 assert_equal.  Returns nothing; `out` must already hold the
// expected digest (as the reference's assert_hash_constraints does with the tree node variable).
static inline Fr hash_native(const HashParams& hp, const Fr& l, const Fr& r) {
    Fr s[3] = {Fr::zero() + hp.ark0[0], l + hp.ark0[1], r + hp.ark0[2]};
    for (int rd = 0; rd < 63; rd++) {
        bool full = rd < 4 || rd >= 59;
        Fr t[3];
        for (int k = 0; k < 3; k++) t[k] = (full || k == 0) ? s[k].pow_u64(5) : s[k];
        for (int j = 0; j < 3; j++) s[j] = hp.mds[j][0] * t[0] + hp.mds[j][1] * t[1] + hp.mds[j][2] * t[2] + hp.rc[rd][j];
    }
    return s[1];
}
static inline void hash_gadget(Composer& cs, const HashParams& hp, uint32_t l, uint32_t r, uint32_t out) {
    uint32_t in[3] = {cs.zero_var, l, r};
    uint32_t s[3];
    for (int k = 0; k < 3; k++) {
        Fr v = cs.var_vals[in[k]] + hp.ark0[k];
        s[k] = cs.add_input(v);
        cs.poly_gate(in[k], cs.zero_var, s[k], Fr::zero(), Fr::one(), Fr::zero(), -Fr::one(), hp.ark0[k], nullptr);
    }
    for (int rd = 0; rd < 63; rd++) {
        bool full = rd < 4 || rd >= 59;
        uint32_t nx[3];
        for (int j = 0; j < 3; j++) {
            Fr selv[5] = {hp.mds[j][0], hp.mds[j][1], hp.mds[j][2], hp.rc[rd][j], -Fr::one()};
            nx[j] = full ? cs.full_affine_transform_gate(s, selv) : cs.partial_affine_transform_gate(s, selv);
        }
        for (int j = 0; j < 3; j++) s[j] = nx[j];
    }
    cs.assert_equal(s[1], out);
}

// Merkle tree with 2^(height-1) leaves => 2^(height-1) - 1 hashes (HEIGHT=4 -> 7, HEIGHT=15 -> 16383).
// n_lookup > 0 appends that many plookup gates against a small XOR table (lookup-enabled variant).
static inline Composer build_merkle_circuit(int height, uint64_t witness_seed, int n_lookup = 0) {
    ensure_init();
    Composer cs;
    SplitMix64 rng(witness_seed);
    cs.prelude(rng);
    HashParams hp(0x504f534549444f4eULL);
    size_t nleaves = (size_t)1 << (height - 1);
    // heap layout: node 0 root, children 2i+1, 2i+2
    size_t nnodes = 2 * nleaves - 1;
    std::vector<Fr> val(nnodes);
    for (size_t i = nleaves - 1; i < nnodes; i++) val[i] = rng.next_fr();
    for (size_t i = nleaves - 1; i-- > 0;) val[i] = hash_native(hp, val[2 * i + 1], val[2 * i + 2]);
    std::vector<uint32_t> var(nnodes);
    for (size_t i = 0; i < nnodes; i++) var[i] = cs.add_input(val[i]);
    for (size_t i = nleaves - 1; i-- > 0;) hash_gadget(cs, hp, var[2 * i + 1], var[2 * i + 2], var[i]);
    if (n_lookup > 0) {
        for (uint64_t a = 0; a < 4; a++)
            for (uint64_t b = 0; b < 4; b++)
                cs.table.push_back({Fr::from_u64(a), Fr::from_u64(b), Fr::from_u64(a ^ b), Fr::from_u64(a + 4 * b)});
        for (int i = 0; i < n_lookup; i++) {
            uint64_t a = rng.next() & 3, b = rng.next() & 3;
            uint32_t va = cs.add_input(Fr::from_u64(a)), vb = cs.add_input(Fr::from_u64(b));
            uint32_t vc = cs.add_input(Fr::from_u64(a ^ b)), vd = cs.add_input(Fr::from_u64(a + 4 * b));
            cs.lookup_gate(va, vb, vc, vd);
        }
    }
    // root == public input: q_l * root + PI = 0 with PI = -root  (merkle-tree/src/constraints.rs:101-108)
    Fr negroot = -val[0];
    cs.poly_gate(var[0], cs.zero_var, cs.zero_var, Fr::zero(), Fr::one(), Fr::zero(), Fr::zero(), Fr::zero(), &negroot);
    return cs;
}

static inline int log2_ceil(size_t x) {
    int l = 0;
    while (((size_t)1 << l) < x) l++;
    return l;
}

struct ProverKeyO {
    int logn;
    size_t n;                                  // padded domain size N
    std::vector<Fr> coeffs[NUM_PK_POLYS];      // N each (not trimmed)
    std::vector<Fr> evals[NUM_PK_POLYS];       // 8N each, coset g*H_8N, natural order
    std::vector<Fr> table[4];                  // N each, padded multisets
    std::vector<Fr> linear_evaluations;        // 8N
    std::vector<Fr> v_h_coset_8n;              // 8N
};

// sigma evaluations on H (permutation/mod.rs:101-166); padded rows map to themselves.
static inline void compute_sigma_evals(const Composer& cs, const Domain& dom, std::vector<Fr> sigma[4]) {
    size_t N = dom.n, n = cs.n();
    size_t nv = cs.var_vals.size();
    std::vector<uint32_t> cnt(nv + 1, 0);
    for (int k = 0; k < 4; k++)
        for (size_t i = 0; i < n; i++) cnt[cs.w[k][i] + 1]++;
    for (size_t v = 0; v < nv; v++) cnt[v + 1] += cnt[v];
    std::vector<uint32_t> pos(cnt.begin(), cnt.end() - 1);
    std::vector<uint64_t> occ(4 * n);  // (gate << 2) | wire, grouped by variable in insertion order
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < 4; k++) occ[pos[cs.w[k][i]]++] = ((uint64_t)i << 2) | k;
    for (int k = 0; k < 4; k++) {
        sigma[k].resize(N);
        Fr kk = K_const(k);
        for (size_t i = 0; i < N; i++) sigma[k][i] = kk * dom.element(i);
    }
    for (size_t v = 0; v < nv; v++) {
        size_t lo = cnt[v], hi = cnt[v + 1];
        for (size_t j = lo; j < hi; j++) {
            uint64_t cur = occ[j], nxt = occ[j + 1 == hi ? lo : j + 1];
            sigma[cur & 3][cur >> 2] = K_const(nxt & 3) * dom.element(nxt >> 2);
        }
    }
}

static inline ProverKeyO preprocess(const Composer& cs) {
    ProverKeyO pk;
    size_t bound = std::max(cs.n(), cs.table.size());
    pk.logn = log2_ceil(bound);
    pk.n = (size_t)1 << pk.logn;
    Domain dom(pk.logn), dom8(pk.logn + 3);
    std::vector<Fr> sigma[4];
    compute_sigma_evals(cs, dom, sigma);
    for (int s = 0; s < NUM_PK_POLYS; s++) {
        std::vector<Fr> ev = s < NUM_SELECTORS ? cs.q[s] : sigma[s - NUM_SELECTORS];
        ev.resize(pk.n, Fr::zero());
        pk.coeffs[s] = dom.ifft(ev);
        pk.evals[s] = dom8.coset_fft(pk.coeffs[s]);
    }
    for (int c = 0; c < 4; c++) {
        std::vector<Fr>& t = pk.table[c];
        for (auto& row : cs.table) t.push_back(row[c]);
        if (t.empty()) t.push_back(Fr::zero());
        t.resize(pk.n, t[0]);  // multiset.rs:70-79
    }
    std::vector<Fr> x = {Fr::zero(), Fr::one()};
    pk.linear_evaluations = dom8.coset_fft(x);
    // preprocess.rs:498-520
    pk.v_h_coset_8n.resize(8 * pk.n);
    Fr cg = fr_generator().pow_u64(pk.n);
    Fr wn = dom8.omega.pow_u64(pk.n);  // 8th root of unity
    Fr p = cg;
    for (size_t i = 0; i < 8 * pk.n; i++) {
        pk.v_h_coset_8n[i] = p - Fr::one();
        p = p * wn;
    }
    return pk;
}

// KZG10 setup with known tau: fixed-base windowed scalar multiplication of the generator.
static inline std::vector<G1Affine> srs_from_tau(const Fr& tau, size_t n) {
    G1 g = G1::from_affine(g1_generator());
    const int WB = 8, NW = 32;
    std::vector<G1> tabj((size_t)NW << WB);
    G1 base = g;
    for (int j = 0; j < NW; j++) {
        tabj[(size_t)j << WB] = G1::infinity();
        for (int d = 1; d < (1 << WB); d++) tabj[((size_t)j << WB) + d] = tabj[((size_t)j << WB) + d - 1].add(base);
        for (int b = 0; b < WB; b++) base = base.dbl();
    }
    std::vector<G1Affine> tab;
    g1_batch_to_affine(tabj, tab);
    std::vector<Fr> pw(n);
    Fr p = Fr::one();
    for (size_t i = 0; i < n; i++) {
        pw[i] = p;
        p = p * tau;
    }
    std::vector<G1> out(n);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) {
        uint64_t k[4];
        pw[i].to_canonical(k);
        G1 acc = G1::infinity();
        for (int j = 0; j < NW; j++) {
            unsigned d = (k[j / 8] >> (8 * (j % 8))) & 0xff;
            if (d) acc = acc.add_affine(tab[((size_t)j << WB) + d]);
        }
        out[i] = acc;
    }
    std::vector<G1Affine> aff;
    g1_batch_to_affine(out, aff);
    return aff;
}

}  // namespace zpo
