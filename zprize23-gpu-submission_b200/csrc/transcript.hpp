// Fiat-Shamir transcript of the prover: Merlin v1.0 on STROBE-128 / Keccak-f[1600], with the
// ark-serialize 0.3 encodings ZK-Garage feeds into it ("Prize 1B/plonk-core/src/transcript.rs":27-50).
// Replaces PNP's host transcript ("…/lib/PLONK/src/transcript/{strobe.cpp:21-171, transcript.cuh:21-73}",
// "…/lib/PLONK/src/serialize.cuh":31-84, "…/transcript/flags.hpp":4-33).  Host only (<1 ms per proof).
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "host_math.hpp"

namespace zp {

class Keccak1600 {
public:
    static void permute(uint8_t state[200]) {
        uint64_t a[25];
        for (int i = 0; i < 25; i++) {
            uint64_t w = 0;
            for (int b = 7; b >= 0; b--) w = (w << 8) | state[8 * i + b];
            a[i] = w;
        }
        uint64_t rc = 1;  // round constants from the degree-8 LFSR of FIPS 202 (no table)
        for (int round = 0; round < 24; round++) {
            uint64_t c[5], d;
            for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
            for (int x = 0; x < 5; x++) {
                d = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
                for (int y = 0; y < 25; y += 5) a[y + x] ^= d;
            }
            // rho + pi, walking the (x, y) -> (y, 2x + 3y) orbit
            int x = 1, y = 0;
            uint64_t cur = a[1];
            for (int t = 0; t < 24; t++) {
                int r = ((t + 1) * (t + 2) / 2) % 64;
                int ny = (2 * x + 3 * y) % 5, nx = y;
                x = nx;
                y = ny;
                uint64_t tmp = a[5 * y + x];
                a[5 * y + x] = rol(cur, r);
                cur = tmp;
            }
            for (int yy = 0; yy < 25; yy += 5) {
                uint64_t row[5];
                for (int xx = 0; xx < 5; xx++) row[xx] = a[yy + xx];
                for (int xx = 0; xx < 5; xx++) a[yy + xx] = row[xx] ^ (~row[(xx + 1) % 5] & row[(xx + 2) % 5]);
            }
            // iota
            uint64_t rcw = 0;
            for (int j = 0; j < 7; j++) {
                if (rc & 1) rcw ^= (uint64_t)1 << ((1 << j) - 1);
                rc = (rc << 1) ^ ((rc >> 7) * 0x71);
                rc &= 0xff;
            }
            a[0] ^= rcw;
        }
        for (int i = 0; i < 25; i++)
            for (int b = 0; b < 8; b++) state[8 * i + b] = (uint8_t)(a[i] >> (8 * b));
    }

private:
    static uint64_t rol(uint64_t v, int r) { return r ? (v << r) | (v >> (64 - r)) : v; }
};

class MerlinTranscript {
public:
    explicit MerlinTranscript(const std::string& label) {
        memset(st_, 0, sizeof(st_));
        st_[0] = 1;
        st_[1] = RATE + 2;
        st_[2] = 1;
        st_[3] = 0;
        st_[4] = 1;
        st_[5] = 96;
        memcpy(st_ + 6, "STROBEv1.0.2", 12);
        Keccak1600::permute(st_);
        pos_ = 0;
        pos_begin_ = 0;
        flags_ = 0;
        static const char proto[] = "Merlin v1.0";
        operate(F_M | F_A, (const uint8_t*)proto, sizeof(proto) - 1, nullptr, false);
        append_message("dom-sep", (const uint8_t*)label.data(), label.size());
    }
    void append_message(const char* label, const uint8_t* msg, size_t len) {
        frame(label, len);
        operate(F_A, msg, len, nullptr, false);
    }
    void challenge_bytes(const char* label, uint8_t* out, size_t len) {
        frame(label, len);
        operate(F_I | F_A | F_C, nullptr, len, out, false);
    }
    // ark-serialize encodings ------------------------------------------------------------
    void append_scalar(const char* label, const host::Fr& x) {
        uint64_t c[4];
        x.to_canonical(c);
        append_message(label, (const uint8_t*)c, 32);
    }
    // compressed G1: 48-byte LE x; bit7 = (y > -y) on canonical integers, bit6 = infinity
    void append_point(const char* label, const host::Fq& x, const host::Fq& y, bool inf) {
        uint8_t b[48];
        memset(b, 0, 48);
        if (inf) {
            b[47] |= 0x40;
        } else {
            uint64_t cx[6], cy[6], cny[6];
            x.to_canonical(cx);
            y.to_canonical(cy);
            y.neg().to_canonical(cny);
            memcpy(b, cx, 48);
            bool greater = false;
            for (int i = 5; i >= 0; i--) {
                if (cy[i] != cny[i]) {
                    greater = cy[i] > cny[i];
                    break;
                }
            }
            if (greater) b[47] |= 0x80;
        }
        append_message(label, b, 48);
    }
    // PublicInputs (BTreeMap<usize,F> of the non-zero entries): u64 count || (u64 pos || Fr)*
    void append_public_inputs(const char* label, const std::vector<std::pair<uint64_t, host::Fr>>& pi) {
        std::vector<uint8_t> b(8 + 40 * pi.size());
        uint64_t cnt = pi.size();
        memcpy(b.data(), &cnt, 8);
        for (size_t i = 0; i < pi.size(); i++) {
            uint64_t c[4];
            pi[i].second.to_canonical(c);
            memcpy(b.data() + 8 + 40 * i, &pi[i].first, 8);
            memcpy(b.data() + 16 + 40 * i, c, 32);
        }
        append_message(label, b.data(), b.size());
    }
    // 31 squeezed bytes as a little-endian integer (< 2^248 < r), converted to Montgomery form
    host::Fr challenge_scalar(const char* label) {
        uint8_t buf[32];
        memset(buf, 0, 32);
        challenge_bytes(label, buf, 31);
        uint64_t c[4];
        memcpy(c, buf, 32);
        return host::Fr::from_canonical(c);
    }

private:
    enum { F_I = 1, F_A = 2, F_C = 4, F_T = 8, F_M = 16, F_K = 32 };
    static const int RATE = 166;
    uint8_t st_[200];
    int pos_, pos_begin_, flags_;

    void frame(const char* label, size_t len) {
        uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
        operate(F_M | F_A, (const uint8_t*)label, strlen(label), nullptr, false);
        operate(F_M | F_A, le, 4, nullptr, true);
    }
    void run_f() {
        st_[pos_] ^= (uint8_t)pos_begin_;
        st_[pos_ + 1] ^= 0x04;
        st_[RATE + 1] ^= 0x80;
        Keccak1600::permute(st_);
        pos_ = 0;
        pos_begin_ = 0;
    }
    void duplex_in(uint8_t byte) {
        st_[pos_++] ^= byte;
        if (pos_ == RATE) run_f();
    }
    // one STROBE operation; `more` continues the previous operation of the same kind
    void operate(int flags, const uint8_t* in, size_t len, uint8_t* out, bool more) {
        if (!more) {
            int old_begin = pos_begin_;
            pos_begin_ = pos_ + 1;
            flags_ = flags;
            duplex_in((uint8_t)old_begin);
            duplex_in((uint8_t)flags);
            if ((flags & (F_C | F_K)) && pos_ != 0) run_f();
        }
        if (out) {
            for (size_t i = 0; i < len; i++) {
                out[i] = st_[pos_];
                st_[pos_] = 0;
                if (++pos_ == RATE) run_f();
            }
        } else {
            for (size_t i = 0; i < len; i++) duplex_in(in[i]);
        }
    }
};

}  // namespace zp
