// ORACLE — TEST INFRASTRUCTURE ONLY.
// Thin extern "C" shim compiled TOGETHER WITH the reference's own, unmodified
// "Prize 1B/plonk-core/lib/PLONK/src/transcript/strobe.cpp" (taken where it lies under /root/reference,
// never copied) into oracle/_ref/libref_strobe.so.  It replays Transcript::new / append_message /
// challenge_bytes exactly as "…/transcript/transcript.cuh":24-37,58-64 does on the reference's
// Strobe128 class, so that oracle/zp_transcript.hpp can be compared against the reference's code.
#include "strobe.h"
#include <string>
#include <cstring>

static std::vector<uint8_t> le32(size_t x) {
    std::vector<uint8_t> b(4);
    for (int i = 0; i < 4; i++) b[i] = (uint8_t)(x >> (8 * i));
    return b;
}
static void append_message(Strobe128& s, const std::string& label, const uint8_t* msg, size_t n) {
    std::vector<uint8_t> l(label.begin(), label.end()), len = le32(n), m(msg, msg + n);
    s.meta_ad(l, false);
    s.meta_ad(len, true);
    s.ad(m, false);
}
static void challenge_bytes(Strobe128& s, const std::string& label, uint8_t* out, size_t n) {
    std::vector<uint8_t> l(label.begin(), label.end()), len = le32(n), d(n, 0);
    s.meta_ad(l, false);
    s.meta_ad(len, true);
    s.prf(d, false);
    memcpy(out, d.data(), n);
}

// Same script encoding as zpo_transcript_script in oracle_capi.cpp.
extern "C" void ref_transcript_script(const char* proto, const uint8_t* script, size_t script_len, uint8_t* out) {
    Strobe128 s = Strobe128::new_instance("Merlin v1.0");
    std::string p(proto);
    append_message(s, "dom-sep", (const uint8_t*)p.data(), p.size());
    size_t pos = 0;
    while (pos < script_len) {
        uint8_t kind = script[pos++];
        uint32_t ll, dl;
        memcpy(&ll, script + pos, 4);
        pos += 4;
        std::string label((const char*)script + pos, ll);
        pos += ll;
        memcpy(&dl, script + pos, 4);
        pos += 4;
        if (kind == 0) {
            append_message(s, label, script + pos, dl);
            pos += dl;
        } else {
            challenge_bytes(s, label, out, dl);
            out += dl;
        }
    }
}
