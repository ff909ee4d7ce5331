// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the
// product path (zprize23-gpu-submission_b200/csrc).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker.
//
// CPU restatement of the BLS12-381 Fr / Fq Montgomery arithmetic that the reference prover uses
// (ark-ff 0.3.0 Fp256/Fp384 on the Rust side; sppark mont_t on the CUDA side,
// "Prize 1B/plonk-core/lib/PLONK/utils/mont/cuda/ff/mont_t.cuh":31-1142; constants
// "…/utils/mont/cpu/ff/bls12-381.hpp":6-67).  Representation matches the FFI: little-endian 64-bit
// limbs in Montgomery form (R = 2^256 for Fr, 2^384 for Fq), always canonical (< p).
//
// Parity status: pinned against the reference's literal constants (tests/test_oracle_constants.py)
// and against the reference's vendored blst (oracle/_ref, tests/test_oracle_vs_ref.py).
#pragma once
#include <cstdint>
#include <cstring>
#include <cassert>
#include <string>
#include <vector>
#include <dlfcn.h>
#include <cstdlib>

namespace zpo {

typedef unsigned __int128 u128;

template <int N>
struct BigN {
    uint64_t v[N];
};

template <int N>
static inline int big_cmp(const uint64_t* a, const uint64_t* b) {
    for (int i = N - 1; i >= 0; i--) {
        if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    }
    return 0;
}
template <int N>
static inline uint64_t big_add(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    u128 c = 0;
    for (int i = 0; i < N; i++) {
        c += (u128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
template <int N>
static inline uint64_t big_sub(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    uint64_t borrow = 0;
    for (int i = 0; i < N; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}

// Field parameters are derived from the modulus alone at start-up (R, R^2, -p^-1 mod 2^64) so that
// the literal constants in the reference can be used as an independent check, not as an input.
template <int N>
struct FieldParams {
    uint64_t p[N];
    uint64_t one[N];  // R mod p
    uint64_t rr[N];   // R^2 mod p
    uint64_t inv;     // -p^-1 mod 2^64
    int bits;

    void init(const uint64_t* modulus, int nbits) {
        memcpy(p, modulus, sizeof(p));
        bits = nbits;
        // Newton iteration for p^-1 mod 2^64
        uint64_t x = 1;
        for (int i = 0; i < 6; i++) x *= 2 - p[0] * x;
        inv = (uint64_t)0 - x;
        // R mod p by 64N modular doublings of 1, R^2 by 64N more.
        uint64_t t[N];
        memset(t, 0, sizeof(t));
        t[0] = 1;
        for (int i = 0; i < 128 * N; i++) {
            uint64_t c = big_add<N>(t, t, t);
            if (c || big_cmp<N>(t, p) >= 0) big_sub<N>(t, t, p);
            if (i == 64 * N - 1) memcpy(one, t, sizeof(t));
        }
        memcpy(rr, t, sizeof(t));
    }
};

// ---- optional accelerator for the CPU baseline: the reference's OWN vendored blst ("Prize 1B/plonk-core/lib/blst",
// compiled where it lies into oracle/_ref/libref_blst.so by build.sh).  `blst_fr_mul` / `blst_fp_mul` are the asm
// Montgomery products and `blst_p1s_mult_pippenger` the CPU MSM the reference links (build.rs:54,61).  Loaded at run
// time from next to liboracle.so; ZPO_NO_BLST=1 or zpo_set_blst(0) selects the plain C++ arithmetic so that the tests
// can compare the two.  All results are canonical field elements / exact group sums, hence byte-identical either way.
struct BlstApi {
    void (*fr_mul)(uint64_t* r, const uint64_t* a, const uint64_t* b) = nullptr;
    void (*fp_mul)(uint64_t* r, const uint64_t* a, const uint64_t* b) = nullptr;
    size_t (*scratch_sizeof)(size_t) = nullptr;
    void (*pippenger)(void* ret, const void* const points[], size_t npoints, const void* const scalars[], size_t nbits,
                      void* scratch) = nullptr;
    void* handle = nullptr;
    bool loaded = false;  // library found with every symbol
    bool on = false;      // currently used
};
static inline BlstApi& blst_api() {
    static BlstApi api;
    return api;
}
static inline void blst_load(const void* addr_in_this_library) {
    BlstApi& a = blst_api();
    if (a.handle) return;
    Dl_info info;
    if (!dladdr(addr_in_this_library, &info) || !info.dli_fname) return;
    std::string path(info.dli_fname);
    size_t slash = path.rfind('/');
    path = (slash == std::string::npos ? std::string(".") : path.substr(0, slash)) + "/_ref/libref_blst.so";
    a.handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!a.handle) return;
    a.fr_mul = (void (*)(uint64_t*, const uint64_t*, const uint64_t*))dlsym(a.handle, "blst_fr_mul");
    a.fp_mul = (void (*)(uint64_t*, const uint64_t*, const uint64_t*))dlsym(a.handle, "blst_fp_mul");
    a.scratch_sizeof = (size_t(*)(size_t))dlsym(a.handle, "blst_p1s_mult_pippenger_scratch_sizeof");
    a.pippenger = (void (*)(void*, const void* const[], size_t, const void* const[], size_t, void*))dlsym(a.handle, "blst_p1s_mult_pippenger");
    a.loaded = a.fr_mul && a.fp_mul && a.scratch_sizeof && a.pippenger;
    const char* off = getenv("ZPO_NO_BLST");
    a.on = a.loaded && !(off && off[0] == '1');
}

template <int N, int TAG>
struct Fp {
    uint64_t v[N];
    static FieldParams<N>& P() {
        static FieldParams<N> params;
        return params;
    }

    static Fp zero() {
        Fp r;
        memset(r.v, 0, sizeof(r.v));
        return r;
    }
    static Fp one() {
        Fp r;
        memcpy(r.v, P().one, sizeof(r.v));
        return r;
    }
    bool is_zero() const {
        uint64_t a = 0;
        for (int i = 0; i < N; i++) a |= v[i];
        return a == 0;
    }
    bool operator==(const Fp& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
    bool operator!=(const Fp& o) const { return !(*this == o); }

    Fp operator+(const Fp& o) const {
        Fp r;
        uint64_t c = big_add<N>(r.v, v, o.v);
        if (c || big_cmp<N>(r.v, P().p) >= 0) big_sub<N>(r.v, r.v, P().p);
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r;
        uint64_t b = big_sub<N>(r.v, v, o.v);
        if (b) big_add<N>(r.v, r.v, P().p);
        return r;
    }
    Fp operator-() const {
        if (is_zero()) return *this;
        Fp r;
        big_sub<N>(r.v, P().p, v);
        return r;
    }
    // CIOS Montgomery product; result canonical.
    Fp operator*(const Fp& o) const {
        if (blst_api().on) {
            Fp r;
            if (N == 4)
                blst_api().fr_mul(r.v, v, o.v);
            else
                blst_api().fp_mul(r.v, v, o.v);
            return r;
        }
        const uint64_t* p = P().p;
        const uint64_t inv = P().inv;
        uint64_t t[N + 2];
        memset(t, 0, sizeof(t));
        for (int i = 0; i < N; i++) {
            u128 c = 0;
            for (int j = 0; j < N; j++) {
                c += (u128)v[j] * o.v[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[N];
            t[N] = (uint64_t)c;
            t[N + 1] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * inv;
            c = (u128)m * p[0] + t[0];
            c >>= 64;
            for (int j = 1; j < N; j++) {
                c += (u128)m * p[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[N];
            t[N - 1] = (uint64_t)c;
            t[N] = t[N + 1] + (uint64_t)(c >> 64);
        }
        Fp r;
        if (t[N] || big_cmp<N>(t, p) >= 0)
            big_sub<N>(r.v, t, p);
        else
            memcpy(r.v, t, sizeof(r.v));
        return r;
    }
    Fp& operator+=(const Fp& o) { return *this = *this + o; }
    Fp& operator-=(const Fp& o) { return *this = *this - o; }
    Fp& operator*=(const Fp& o) { return *this = *this * o; }
    Fp square() const { return *this * *this; }
    Fp dbl() const { return *this + *this; }

    Fp pow(const uint64_t* e, int nlimbs) const {
        Fp r = one();
        bool started = false;
        for (int i = nlimbs * 64 - 1; i >= 0; i--) {
            if (started) r = r.square();
            if ((e[i / 64] >> (i % 64)) & 1) {
                r = started ? r * *this : *this;
                started = true;
            }
        }
        return r;
    }
    Fp pow_u64(uint64_t e) const { return pow(&e, 1); }
    // Fermat inverse; inverse of zero is zero (callers assert as the reference does).
    Fp inverse() const {
        uint64_t e[N];
        uint64_t two[N];
        memset(two, 0, sizeof(two));
        two[0] = 2;
        big_sub<N>(e, P().p, two);
        return pow(e, N);
    }

    // Montgomery <-> canonical
    static Fp from_canonical(const uint64_t* c) {
        Fp a, rr;
        memcpy(a.v, c, sizeof(a.v));
        memcpy(rr.v, P().rr, sizeof(rr.v));
        return a * rr;
    }
    void to_canonical(uint64_t* out) const {
        Fp o;
        memset(o.v, 0, sizeof(o.v));
        o.v[0] = 1;
        Fp r = *this * o;
        memcpy(out, r.v, sizeof(r.v));
    }
    static Fp from_u64(uint64_t x) {
        uint64_t c[N];
        memset(c, 0, sizeof(c));
        c[0] = x;
        return from_canonical(c);
    }
    // canonical little-endian bytes (ark-serialize layout for a field element)
    void to_bytes_le(uint8_t* out) const {
        uint64_t c[N];
        to_canonical(c);
        memcpy(out, c, 8 * N);
    }
    // lexicographic comparison of canonical integers (ark-ff Ord)
    static int cmp_canonical(const Fp& a, const Fp& b) {
        uint64_t x[N], y[N];
        a.to_canonical(x);
        b.to_canonical(y);
        return big_cmp<N>(x, y);
    }
    static Fp from_hex(const char* hex) {  // big-endian hex string, canonical value
        uint64_t c[N];
        memset(c, 0, sizeof(c));
        size_t len = strlen(hex);
        for (size_t i = 0; i < len; i++) {
            char ch = hex[len - 1 - i];
            uint64_t d = (ch >= '0' && ch <= '9') ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : ch - 'A' + 10;
            c[i / 16] |= d << (4 * (i % 16));
        }
        return from_canonical(c);
    }
};

typedef Fp<4, 0> Fr;
typedef Fp<6, 1> Fq;

// BLS12-381 moduli (same literals as "…/utils/mont/cpu/ff/bls12-381.hpp":7-11 and :34-40).
static const uint64_t FR_MODULUS[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                                       0x73eda753299d7d48ULL};
static const uint64_t FQ_MODULUS[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                       0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};

struct FieldInit {
    FieldInit() {
        Fr::P().init(FR_MODULUS, 255);
        Fq::P().init(FQ_MODULUS, 381);
        blst_load((const void*)FR_MODULUS);
    }
};
static inline void ensure_init() { static FieldInit once; (void)once; }

// Fr specifics used by the prover (ark-ff FftParameters of ark-bls12-381 0.3.0;
// in-tree copy: "Prize 1B/plonk-core/lib/PLONK/src/bls12_381/fr.cuh":13,39-53).
static const int FR_TWO_ADICITY = 32;
static inline Fr fr_generator() { return Fr::from_u64(7); }
// TWO_ADIC_ROOT_OF_UNITY = 7^((r-1)/2^32)
static inline Fr fr_two_adic_root() {
    static Fr root;
    static bool done = false;
    if (!done) {
        uint64_t e[4];
        uint64_t pm1[4];
        uint64_t onev[4] = {1, 0, 0, 0};
        big_sub<4>(pm1, FR_MODULUS, onev);
        // shift right by 32
        for (int i = 0; i < 4; i++) e[i] = (pm1[i] >> 32) | (i < 3 ? (pm1[i + 1] << 32) : 0);
        root = fr_generator().pow(e, 4);
        done = true;
    }
    return root;
}
// generator of the size-2^logn subgroup: root^(2^(32-logn)) ("…/PLONK/src/domain.cu":29-36)
static inline Fr fr_root_of_unity(int logn) {
    Fr w = fr_two_adic_root();
    for (int i = logn; i < FR_TWO_ADICITY; i++) w = w.square();
    return w;
}

// Deterministic test RNG (SplitMix64), SURVEY §8d seeds.
struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    Fr next_fr() {
        uint64_t c[4];
        for (int i = 0; i < 4; i++) c[i] = next();
        c[3] &= 0x3fffffffffffffffULL;  // < 2^254 < r
        return Fr::from_canonical(c);
    }
};

}  // namespace zpo
