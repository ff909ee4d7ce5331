// ORACLE — TEST INFRASTRUCTURE ONLY (see zp_field.hpp header).
//
// Merlin v1.0 transcript over STROBE-128 / Keccak-f[1600] plus the ark-serialize 0.3.0 encodings the
// prover feeds into it.  Restates merlin 3.0.0 (pinned in "Prize 1B/Cargo.lock") and
// "Prize 1B/plonk-core/src/transcript.rs":27-50.  The reference tree carries its own C++ restatement
// ("…/lib/PLONK/src/transcript/strobe.cpp":21-171, "…/transcript.cuh":21-73, "…/serialize.cuh":31-84,
// "…/flags.hpp":4-33) which oracle/_ref compiles and tests/test_oracle_vs_ref.py replays against this
// file; the Merlin `equivalence_simple` known-answer vector is checked in tests/test_oracle_transcript.py.
#pragma once
#include "zp_curve.hpp"

namespace zpo {

static inline uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

static inline void keccak_f1600(uint64_t st[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5];
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            uint64_t t = bc[(i + 4) % 5] ^ rotl64(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        uint64_t t = st[1];
        for (int i = 0; i < 24; i++) {
            int j = PIL[i];
            uint64_t b = st[j];
            st[j] = rotl64(t, ROT[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= RC[round];
    }
}

struct Strobe128 {
    static const int R = 166;
    enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
    uint8_t st[200];
    int pos, pos_begin, cur_flags;

    explicit Strobe128(const std::string& protocol_label) {
        memset(st, 0, sizeof(st));
        const uint8_t hdr[6] = {1, R + 2, 1, 0, 1, 96};
        memcpy(st, hdr, 6);
        memcpy(st + 6, "STROBEv1.0.2", 12);
        permute();
        pos = pos_begin = cur_flags = 0;
        meta_ad((const uint8_t*)protocol_label.data(), protocol_label.size(), false);
    }
    void permute() {
        uint64_t w[25];
        memcpy(w, st, 200);  // little-endian host
        keccak_f1600(w);
        memcpy(st, w, 200);
    }
    void run_f() {
        st[pos] ^= (uint8_t)pos_begin;
        st[pos + 1] ^= 0x04;
        st[R + 1] ^= 0x80;
        permute();
        pos = 0;
        pos_begin = 0;
    }
    void absorb(const uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; i++) {
            st[pos] ^= d[i];
            if (++pos == R) run_f();
        }
    }
    void squeeze(uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; i++) {
            d[i] = st[pos];
            st[pos] = 0;
            if (++pos == R) run_f();
        }
    }
    void begin_op(int flags, bool more) {
        if (more) {
            assert(cur_flags == flags);
            return;
        }
        assert((flags & FLAG_T) == 0);
        uint8_t old_begin = (uint8_t)pos_begin;
        pos_begin = pos + 1;
        cur_flags = flags;
        uint8_t b[2] = {old_begin, (uint8_t)flags};
        absorb(b, 2);
        bool force_f = (flags & (FLAG_C | FLAG_K)) != 0;
        if (force_f && pos != 0) run_f();
    }
    void meta_ad(const uint8_t* d, size_t n, bool more) {
        begin_op(FLAG_M | FLAG_A, more);
        absorb(d, n);
    }
    void ad(const uint8_t* d, size_t n, bool more) {
        begin_op(FLAG_A, more);
        absorb(d, n);
    }
    void prf(uint8_t* d, size_t n, bool more) {
        begin_op(FLAG_I | FLAG_A | FLAG_C, more);
        squeeze(d, n);
    }
};

// ark-serialize: compressed G1 = 48-byte LE x, bit7 of last byte = (y > -y), bit6 = infinity.
static inline void serialize_g1(const G1Affine& p, uint8_t out[48]) {
    if (p.inf) {
        memset(out, 0, 48);
        out[47] |= 1 << 6;
        return;
    }
    p.x.to_bytes_le(out);
    Fq ny = -p.y;
    if (Fq::cmp_canonical(p.y, ny) > 0) out[47] |= 1 << 7;
}

struct Transcript {
    Strobe128 strobe;
    explicit Transcript(const std::string& label) : strobe("Merlin v1.0") {
        append_message("dom-sep", (const uint8_t*)label.data(), label.size());
    }
    void append_message(const char* label, const uint8_t* msg, size_t n) {
        uint32_t len = (uint32_t)n;
        uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
        strobe.meta_ad((const uint8_t*)label, strlen(label), false);
        strobe.meta_ad(le, 4, true);
        strobe.ad(msg, n, false);
    }
    void challenge_bytes(const char* label, uint8_t* out, size_t n) {
        uint32_t len = (uint32_t)n;
        uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
        strobe.meta_ad((const uint8_t*)label, strlen(label), false);
        strobe.meta_ad(le, 4, true);
        strobe.prf(out, n, false);
    }
    void append_fr(const char* label, const Fr& x) {
        uint8_t b[32];
        x.to_bytes_le(b);
        append_message(label, b, 32);
    }
    void append_g1(const char* label, const G1Affine& p) {
        uint8_t b[48];
        serialize_g1(p, b);
        append_message(label, b, 48);
    }
    // PublicInputs = BTreeMap<usize, F> of the non-zero entries: u64 count || (u64 pos || 32-byte Fr)...
    // ("Prize 1B/plonk-core/src/proof_system/pi.rs":28-37,55-62)
    void append_pi(const char* label, const std::vector<std::pair<uint64_t, Fr>>& pi) {
        std::vector<uint8_t> b(8 + 40 * pi.size());
        uint64_t cnt = pi.size();
        memcpy(b.data(), &cnt, 8);
        for (size_t i = 0; i < pi.size(); i++) {
            memcpy(b.data() + 8 + 40 * i, &pi[i].first, 8);
            pi[i].second.to_bytes_le(b.data() + 16 + 40 * i);
        }
        append_message(label, b.data(), b.size());
    }
    // transcript.rs:33-44: 31 squeezed bytes, little-endian integer, to Montgomery
    Fr challenge_scalar(const char* label) {
        uint8_t buf[32];
        memset(buf, 0, 32);
        challenge_bytes(label, buf, 31);
        uint64_t c[4];
        memcpy(c, buf, 32);
        return Fr::from_canonical(c);
    }
};

}  // namespace zpo
