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
    // Permutation::variable_map (permutation/mod.rs:25-98) flattened: (variable, (gate << 2) | wire) in the order the
    // reference calls add_variable_to_map — the cycle of a variable follows this insertion order (mod.rs:101-137).
    // Wire cells the reference never maps (range.rs:185-187) are simply absent: sigma is the identity there.
    std::vector<std::pair<uint32_t, uint32_t>> perm_log;

    size_t n() const { return w[0].size(); }
    uint32_t add_input(const Fr& v) {
        var_vals.push_back(v);
        return (uint32_t)var_vals.size() - 1;
    }
    void map_wire(uint32_t var, int wire, size_t gate) { perm_log.push_back({var, (uint32_t)((gate << 2) | (uint32_t)wire)}); }
    void push_selector_row() {
        for (int s = 0; s < NUM_SELECTORS; s++) q[s].push_back(Fr::zero());
    }
    // add_variables_to_map order: left, right, output, fourth (permutation/mod.rs:68-88)
    void push_gate(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        size_t g = n();
        w[0].push_back(a);
        w[1].push_back(b);
        w[2].push_back(c);
        w[3].push_back(d);
        map_wire(a, 0, g);
        map_wire(b, 1, g);
        map_wire(c, 2, g);
        map_wire(d, 3, g);
        push_selector_row();
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
    // composer.rs:334-351
    void constrain_to_constant(uint32_t a, const Fr& constant) {
        poly_gate(a, a, a, Fr::zero(), Fr::one(), Fr::zero(), Fr::zero(), -constant, nullptr);
    }
    // arithmetic.rs:97-164 — fan-in-3 arithmetic gate; the output wire is computed when not supplied
    uint32_t arithmetic_gate(uint32_t a, uint32_t b, const uint32_t* c_in, Fr qm, Fr ql, Fr qr, Fr qo, Fr qc, Fr q4, uint32_t w4) {
        uint32_t c;
        if (c_in)
            c = *c_in;
        else
            c = add_input((qm * var_vals[a] * var_vals[b] + ql * var_vals[a] + qr * var_vals[b] + qc + q4 * var_vals[w4]) * (-qo));
        push_gate(a, b, c, w4);
        sel(Q_M) = qm;
        sel(Q_L) = ql;
        sel(Q_R) = qr;
        sel(Q_O) = qo;
        sel(Q_C) = qc;
        sel(Q_4) = q4;
        sel(Q_ARITH) = Fr::one();
        return c;
    }
    static std::vector<uint8_t> bits_le(const Fr& v) {  // into_repr().to_bits_le()
        uint64_t c[4];
        v.to_canonical(c);
        std::vector<uint8_t> b(256);
        for (int i = 0; i < 256; i++) b[i] = (c[i >> 6] >> (i & 63)) & 1;
        return b;
    }

    // range.rs:27-211 — quads of the witness accumulated base 4, four accumulators per gate laid out d, c, b, a
    void range_gate(uint32_t witness, size_t num_bits) {
        assert(num_bits % 2 == 0);
        const size_t n0 = n();
        std::vector<uint8_t> bits = bits_le(var_vals[witness]);
        size_t num_gates = num_bits >> 3;
        if (num_bits % 8 != 0) num_gates++;
        const size_t num_quads = num_gates * 4;
        const size_t pad = 1 + (((num_quads << 1) - num_bits) >> 1);
        const size_t used_gates = num_gates + 1;
        auto add_wire = [&](size_t i, uint32_t var) {
            const size_t gate = n0 + i / 4;
            static const int wire_of[4] = {3, 2, 1, 0};  // i % 4: fourth, output, right, left
            const int k = wire_of[i % 4];
            assert(w[k].size() == gate);
            w[k].push_back(var);
            map_wire(var, k, gate);
        };
        std::vector<uint32_t> accumulators;
        Fr acc = Fr::zero(), four = Fr::from_u64(4);
        for (size_t i = 0; i < pad; i++) add_wire(i, zero_var);
        for (size_t i = pad; i <= num_quads; i++) {
            size_t bit_index = (num_quads - i) << 1;
            uint64_t quad = bits[bit_index] + 2 * bits[bit_index + 1];
            acc = four * acc + Fr::from_u64(quad);
            uint32_t v = add_input(acc);
            accumulators.push_back(v);
            add_wire(i, v);
        }
        for (size_t g = 0; g < used_gates; g++) {
            push_selector_row();
            sel(Q_RANGE) = Fr::one();
        }
        sel(Q_RANGE) = Fr::zero();  // last gate: holds one quad in the fourth wire only
        // left, right and output of the last gate are pushed WITHOUT a permutation entry (range.rs:185-187)
        w[0].push_back(zero_var);
        w[1].push_back(zero_var);
        w[2].push_back(zero_var);
        assert(w[0].size() == n0 + used_gates && w[3].size() == n0 + used_gates);
        assert_equal(accumulators.back(), witness);
    }

    // logic.rs:36-326 — returns the variable holding a (xor | and) b over num_bits
    uint32_t logic_gate(uint32_t a, uint32_t b, size_t num_bits, bool is_xor) {
        assert((num_bits & 1) == 0);
        const size_t num_quads = num_bits >> 1;
        std::vector<uint8_t> al = bits_le(var_vals[a]), bl = bits_le(var_vals[b]);
        // to_bits_be().skip(256 - num_bits): index j of that slice is bit (num_bits - 1 - j)
        auto be = [&](const std::vector<uint8_t>& le, size_t j) { return le[num_bits - 1 - j]; };
        Fr left = Fr::zero(), right = Fr::zero(), out = Fr::zero(), four = Fr::from_u64(4);
        size_t g = n();
        map_wire(zero_var, 0, g);
        map_wire(zero_var, 1, g);
        map_wire(zero_var, 3, g);
        w[0].push_back(zero_var);
        w[1].push_back(zero_var);
        w[3].push_back(zero_var);
        g++;
        for (size_t i = 0; i < num_quads; i++) {
            uint8_t lq = (uint8_t)((be(al, 2 * i) << 1) + be(al, 2 * i + 1));
            uint8_t rq = (uint8_t)((be(bl, 2 * i) << 1) + be(bl, 2 * i + 1));
            uint8_t oq = is_xor ? (lq ^ rq) : (lq & rq);
            left = left * four + Fr::from_u64(lq);
            right = right * four + Fr::from_u64(rq);
            out = out * four + Fr::from_u64(oq);
            uint32_t va = add_input(left), vb = add_input(right), vc = add_input(Fr::from_u64((uint64_t)lq * rq)), v4 = add_input(out);
            map_wire(va, 0, g);
            map_wire(vb, 1, g);
            map_wire(v4, 3, g);
            map_wire(vc, 2, g - 1);
            w[0].push_back(va);
            w[1].push_back(vb);
            w[2].push_back(vc);
            w[3].push_back(v4);
            g++;
        }
        map_wire(zero_var, 2, g - 1);
        w[2].push_back(zero_var);
        for (size_t i = 0; i < num_quads; i++) {
            push_selector_row();
            sel(Q_C) = is_xor ? -Fr::one() : Fr::one();
            sel(Q_LOGIC) = is_xor ? -Fr::one() : Fr::one();
        }
        push_selector_row();  // last row: no-op
        assert(w[0].size() == g && w[2].size() == g && q[0].size() == g);
        return w[3].back();
    }
};

// ---- JubJub (ed-on-bls12-381): -x^2 + y^2 = 1 + d x^2 y^2 over Fr, d = -(10240/10241)
// (Montgomery literals of a, d: "…/lib/PLONK/src/bls12_381/edwards.cu":5-31; ark-ed-on-bls12-381 0.3 is not vendored —
// the generator below is that crate's AFFINE_GENERATOR_COEFFS, checked to be on the curve by tests/test_oracle_gates.py)
struct TEPoint {
    Fr x, y;
};
static inline Fr te_d() { return -(Fr::from_u64(10240) * Fr::from_u64(10241).inverse()); }
static inline TEPoint te_identity() { return {Fr::zero(), Fr::one()}; }
static inline TEPoint te_neg(const TEPoint& p) { return {-p.x, p.y}; }
static inline bool te_on_curve(const TEPoint& p) {
    Fr x2 = p.x.square(), y2 = p.y.square();
    return (y2 - x2) == Fr::one() + te_d() * x2 * y2;
}
static inline TEPoint te_add(const TEPoint& p, const TEPoint& q) {  // a = -1
    Fr x1y2 = p.x * q.y, y1x2 = p.y * q.x, y1y2 = p.y * q.y, x1x2 = p.x * q.x;
    Fr k = te_d() * x1y2 * y1x2;
    return {(x1y2 + y1x2) * (Fr::one() + k).inverse(), (y1y2 + x1x2) * (Fr::one() - k).inverse()};
}
static inline TEPoint te_generator() {
    static const char* GX = "11dafe5d23e1218086a365b99fbf3d3be72f6afd7d1f72623e6b071492d1122b";
    static const char* GY = "1d523cf1ddab1a1793132e78c866c0c33e26ba5cc220fed7cc3f870e59d292aa";
    return {Fr::from_hex(GX), Fr::from_hex(GY)};
}
// width-2 NAF of the canonical value, least significant digit first (ark-ff BigInteger::find_wnaf, w = 2)
static inline std::vector<int> wnaf2(const Fr& s) {
    uint64_t e[5] = {0, 0, 0, 0, 0};
    s.to_canonical(e);
    std::vector<int> res;
    auto is_zero = [&] { return !(e[0] | e[1] | e[2] | e[3] | e[4]); };
    while (!is_zero()) {
        int z = 0;
        if (e[0] & 1) {
            z = 2 - (int)(e[0] & 3);  // 1 -> 1, 3 -> -1
            if (z > 0) {
                e[0] -= 1;  // odd: no borrow
            } else {
                for (int i = 0; i < 5; i++)
                    if (++e[i]) break;
            }
        }
        res.push_back(z);
        for (int i = 0; i < 4; i++) e[i] = (e[i] >> 1) | (e[i + 1] << 63);
        e[4] >>= 1;
    }
    return res;
}

// ecc/scalar_mul/fixed_base.rs:52-163 (gates: ecc/curve_addition/fixed_base_gate.rs:78-112)
static inline void fixed_base_scalar_mul(Composer& cs, uint32_t scalar, const TEPoint& base, uint32_t out_xy[2]) {
    const size_t num_bits = 255;  // Fr::MODULUS_BITS
    std::vector<TEPoint> multiples(num_bits);
    multiples[0] = base;
    for (size_t i = 1; i < num_bits; i++) multiples[i] = te_add(multiples[i - 1], multiples[i - 1]);
    std::reverse(multiples.begin(), multiples.end());
    std::vector<int> wnaf = wnaf2(cs.var_vals[scalar]);
    assert(wnaf.size() <= num_bits);
    std::vector<Fr> scalar_acc{Fr::zero()};
    std::vector<TEPoint> point_acc{te_identity()};
    std::vector<Fr> xy_alphas;
    const size_t tz = num_bits - wnaf.size();
    for (size_t i = 0; i < tz; i++) {
        scalar_acc.push_back(Fr::zero());
        point_acc.push_back(te_identity());
        xy_alphas.push_back(Fr::zero());
    }
    for (size_t i = 0; i < wnaf.size(); i++) {
        const int entry = wnaf[wnaf.size() - 1 - i];
        const size_t index = i + tz;
        Fr s_add = Fr::zero();
        TEPoint p_add = te_identity();
        if (entry == 1) {
            s_add = Fr::one();
            p_add = multiples[index];
        } else if (entry == -1) {
            s_add = -Fr::one();
            p_add = te_neg(multiples[index]);
        }
        scalar_acc.push_back(scalar_acc[index].dbl() + s_add);
        point_acc.push_back(te_add(point_acc[index], p_add));
        xy_alphas.push_back(p_add.x * p_add.y);
    }
    for (size_t i = 0; i < num_bits; i++) {
        uint32_t acc_x = cs.add_input(point_acc[i].x), acc_y = cs.add_input(point_acc[i].y);
        uint32_t acc_bit = cs.add_input(scalar_acc[i]);
        if (i == 0) {
            cs.constrain_to_constant(acc_x, Fr::zero());
            cs.constrain_to_constant(acc_y, Fr::one());
            cs.constrain_to_constant(acc_bit, Fr::zero());
        }
        uint32_t xy_alpha = cs.add_input(xy_alphas[i]);
        cs.push_gate(acc_x, acc_y, xy_alpha, acc_bit);
        cs.sel(Q_L) = multiples[i].x;
        cs.sel(Q_R) = multiples[i].y;
        cs.sel(Q_C) = multiples[i].x * multiples[i].y;
        cs.sel(Q_FIXED) = Fr::one();
    }
    uint32_t acc_x = cs.add_input(point_acc[num_bits].x), acc_y = cs.add_input(point_acc[num_bits].y);
    uint32_t last_bit = cs.add_input(scalar_acc[num_bits]);
    uint32_t zv = cs.zero_var;
    cs.arithmetic_gate(acc_x, acc_y, &zv, Fr::zero(), Fr::zero(), Fr::zero(), Fr::zero(), Fr::zero(), Fr::zero(), last_bit);
    cs.assert_equal(last_bit, scalar);
    out_xy[0] = acc_x;
    out_xy[1] = acc_y;
}

// ecc/curve_addition/variable_base_gate.rs:25-98
static inline void point_addition_gate(Composer& cs, const uint32_t a[2], const uint32_t b[2], uint32_t out_xy[2]) {
    TEPoint p1{cs.var_vals[a[0]], cs.var_vals[a[1]]}, p2{cs.var_vals[b[0]], cs.var_vals[b[1]]};
    TEPoint p3 = te_add(p1, p2);
    uint32_t x1y2 = cs.add_input(p1.x * p2.y), x3 = cs.add_input(p3.x), y3 = cs.add_input(p3.y);
    cs.push_gate(a[0], a[1], b[0], b[1]);
    cs.sel(Q_VAR) = Fr::one();
    cs.push_gate(x3, y3, cs.zero_var, x1y2);
    out_xy[0] = x3;
    out_xy[1] = y3;
}

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

// Circuit kinds of the oracle front end (zpo_ctx_new_kind):
//   0 Poseidon-Merkle (above)
//   1 every TurboPLONK widget at once: range, logic (xor + and), fixed-base scalar multiplication, curve addition,
//     multiplication gates (q_m != 0), a few Poseidon rounds and optional plookup rows — inputs of the reference's own
//     gadget tests (range.rs:213-250, logic.rs:353-400, ecc/scalar_mul/fixed_base.rs:176-215)
//   2 no constant selector anywhere (q_c == 0): prelude + multiplications + equality + public input
//   3 no arithmetic selector anywhere (q_arith == 0): blinding rows + range gadget cells only, no public input
// reps > 1 (kind 1 only) repeats the gadget block that many times with fresh random operands, to reach domain sizes where
// the prover takes its production routes (precomputed-table MSM, batch-affine rounds, 3-pass NTTs)
static inline Composer build_custom_circuit(int kind, uint64_t witness_seed, int n_lookup = 0, int reps = 1) {
    ensure_init();
    Composer cs;
    SplitMix64 rng(witness_seed);
    if (kind == 3) {
        // hand-made: the reference's prelude and assert_equal are arithmetic gates, so this shape only exists at the FFI
        cs.zero_var = cs.add_input(Fr::zero());
        for (int i = 0; i < 3; i++) {
            uint32_t r1 = cs.add_input(rng.next_fr()), r2 = cs.add_input(rng.next_fr());
            cs.push_gate(r1, r2, cs.add_input(rng.next_fr()), cs.add_input(rng.next_fr()));
        }
        for (int rep = 0; rep < 5; rep++) {
            // four quads per row, d_next = 4a + quad: 12 rows of a running base-4 accumulator
            Fr acc = Fr::zero();
            uint32_t prev = cs.zero_var;
            for (int row = 0; row < 12; row++) {
                uint32_t v[4];
                for (int k = 0; k < 4; k++) {
                    acc = acc * Fr::from_u64(4) + Fr::from_u64(rng.next() & 3);
                    v[k] = cs.add_input(acc);
                }
                cs.push_gate(v[2], v[1], v[0], prev);  // d (prev), c, b, a ascending
                cs.sel(Q_RANGE) = Fr::one();
                prev = v[3];
            }
            cs.push_gate(cs.zero_var, cs.zero_var, cs.zero_var, prev);
        }
        return cs;
    }
    cs.prelude(rng);
    if (kind == 2) {
        uint32_t x = cs.add_input(rng.next_fr());
        for (int i = 0; i < 40; i++) {
            uint32_t y = cs.add_input(rng.next_fr());
            uint32_t z = cs.arithmetic_gate(x, y, nullptr, Fr::one(), Fr::zero(), Fr::zero(), -Fr::one(), Fr::zero(), Fr::zero(), cs.zero_var);
            uint32_t zc = cs.add_input(cs.var_vals[z]);
            cs.assert_equal(z, zc);
            x = zc;
        }
        Fr neg = -cs.var_vals[x];
        cs.poly_gate(x, cs.zero_var, cs.zero_var, Fr::zero(), Fr::one(), Fr::zero(), Fr::zero(), Fr::zero(), &neg);
        return cs;
    }
    // ---- kind 1
    uint32_t digest = 0;
    for (int rep = 0; rep < (reps < 1 ? 1 : reps); rep++) {
    // range gadget: 34 bits (padding case) and 32 bits (genesis-quad case)
    cs.range_gate(cs.add_input(Fr::from_u64(rep ? (rng.next() & ((((uint64_t)1) << 34) - 1)) : ((uint64_t)1 << 34) - 1)), 34);
    cs.range_gate(cs.add_input(Fr::from_u64(rng.next() & 0xffffffffULL)), 32);
    // logic gadget
    uint32_t xr = cs.logic_gate(cs.add_input(Fr::from_u64(500)), cs.add_input(Fr::from_u64(357)), 10, true);
    cs.constrain_to_constant(xr, Fr::from_u64(500 ^ 357));
    uint32_t ar = cs.logic_gate(cs.add_input(Fr::from_u64(469)), cs.add_input(Fr::from_u64(321)), 10, false);
    cs.constrain_to_constant(ar, Fr::from_u64(469 & 321));
    uint64_t la = rng.next(), lb = rng.next();
    uint32_t xr64 = cs.logic_gate(cs.add_input(Fr::from_u64(la)), cs.add_input(Fr::from_u64(lb)), 64, true);
    cs.constrain_to_constant(xr64, Fr::from_u64(la ^ lb));
    // fixed-base scalar multiplication of the JubJub generator by the reference test's scalar, and by a random one
    static const uint8_t SC[32] = {182, 44, 247, 214, 94, 14, 151, 208, 130, 16, 200, 204, 147, 32, 104, 166,
                                   0, 59, 52, 1, 1, 59, 103, 6, 169, 175, 51, 101, 234, 180, 125, 4};
    uint64_t sc[4];
    memcpy(sc, SC, 32);
    if (rep) {
        sc[0] ^= rng.next();
        sc[1] ^= rng.next();
    }
    uint32_t p1[2], p2[2], p3[2];
    fixed_base_scalar_mul(cs, cs.add_input(Fr::from_canonical(sc)), te_generator(), p1);
    uint64_t s2[4] = {rng.next(), rng.next(), rng.next(), rng.next() >> 6};
    fixed_base_scalar_mul(cs, cs.add_input(Fr::from_canonical(s2)), te_generator(), p2);
    // curve addition of the two results, then of the sum with itself
    point_addition_gate(cs, p1, p2, p3);
    uint32_t p4[2];
    point_addition_gate(cs, p3, p3, p4);
    // multiplication gates
    uint32_t x = cs.add_input(rng.next_fr());
    for (int i = 0; i < 16; i++) {
        uint32_t y = cs.add_input(rng.next_fr());
        x = cs.arithmetic_gate(x, y, nullptr, rng.next_fr(), rng.next_fr(), rng.next_fr(), -Fr::one(), rng.next_fr(), rng.next_fr(),
                               cs.add_input(rng.next_fr()));
    }
    // one Poseidon-shaped hash
    HashParams hp(0x504f534549444f4eULL);
    Fr l = rng.next_fr(), r = rng.next_fr();
    digest = cs.add_input(hash_native(hp, l, r));
    hash_gadget(cs, hp, cs.add_input(l), cs.add_input(r), digest);
    }  // reps
    if (n_lookup > 0) {
        for (uint64_t a = 0; a < 4; a++)
            for (uint64_t b = 0; b < 4; b++)
                cs.table.push_back({Fr::from_u64(a), Fr::from_u64(b), Fr::from_u64(a ^ b), Fr::from_u64(a + 4 * b)});
        for (int i = 0; i < n_lookup; i++) {
            uint64_t a = rng.next() & 3, b = rng.next() & 3;
            cs.lookup_gate(cs.add_input(Fr::from_u64(a)), cs.add_input(Fr::from_u64(b)), cs.add_input(Fr::from_u64(a ^ b)),
                           cs.add_input(Fr::from_u64(a + 4 * b)));
        }
    }
    Fr neg = -cs.var_vals[digest];
    cs.poly_gate(digest, cs.zero_var, cs.zero_var, Fr::zero(), Fr::one(), Fr::zero(), Fr::zero(), Fr::zero(), &neg);
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

// sigma evaluations on H (permutation/mod.rs:101-166); cells without a permutation entry (padding rows, the unmapped
// cells of a range gadget) map to themselves.  The cycle of a variable visits its cells in insertion order.
static inline void compute_sigma_evals(const Composer& cs, const Domain& dom, std::vector<Fr> sigma[4]) {
    size_t N = dom.n;
    size_t nv = cs.var_vals.size(), m = cs.perm_log.size();
    std::vector<uint32_t> cnt(nv + 1, 0);
    for (auto& e : cs.perm_log) cnt[e.first + 1]++;
    for (size_t v = 0; v < nv; v++) cnt[v + 1] += cnt[v];
    std::vector<uint32_t> pos(cnt.begin(), cnt.end() - 1);
    std::vector<uint32_t> occ(m);  // (gate << 2) | wire, grouped by variable, insertion order kept (stable)
    for (auto& e : cs.perm_log) occ[pos[e.first]++] = e.second;
    for (int k = 0; k < 4; k++) {
        sigma[k].resize(N);
        Fr kk = K_const(k);
        for (size_t i = 0; i < N; i++) sigma[k][i] = kk * dom.element(i);
    }
    for (size_t v = 0; v < nv; v++) {
        size_t lo = cnt[v], hi = cnt[v + 1];
        for (size_t j = lo; j < hi; j++) {
            uint32_t cur = occ[j], nxt = occ[j + 1 == hi ? lo : j + 1];
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
        bool all_zero = true;
        for (size_t i = 0; i < ev.size() && all_zero; i++) all_zero = ev[i].is_zero();
        if (all_zero) {  // the transform of the zero vector, without running it
            pk.coeffs[s].assign(pk.n, Fr::zero());
            pk.evals[s].assign(8 * pk.n, Fr::zero());
            continue;
        }
        pk.coeffs[s] = dom.ifft(ev);
        pk.evals[s] = dom8.coset_fft(pk.coeffs[s]);
    }
    for (int c = 0; c < 4; c++) {
        std::vector<Fr>& t = pk.table[c];
        for (auto& row : cs.table) t.push_back(row[c]);
        if (t.empty()) t.push_back(Fr::zero());
        t.resize(pk.n, t[0]);  // multiset.rs:70-79
    }
    // coset_fft of the polynomial X: the coset points g * omega_8N^i themselves (preprocess.rs:281-284)
    pk.linear_evaluations.resize(8 * pk.n);
    {
        const Fr g = fr_generator();
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)(8 * pk.n); i++) pk.linear_evaluations[i] = g * dom8.element(i);
    }
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
