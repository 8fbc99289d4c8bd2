#include "proto.h"
#include <cstring>

namespace amps {

void expandbits(uint8_t *out, int nbits, uint64_t val) {
    for (int i = 0; i < nbits; ++i) out[i] = (uint8_t)((val >> (nbits - 1 - i)) & 1u);
}

// (m(x) x^12) mod g(x), g = x^12+x^10+x^8+x^5+x^4+x^3+1: what itpp::BCH(63,2,true).encode appends
// after the information bits (lib/focc_impl.cc:156-176 pads 23 zeros and keeps the last 40 bits)
std::array<uint8_t, 40> bch_encode_40_28(const uint8_t *info28) {
    std::array<uint8_t, 40> out{};
    uint32_t rem = 0;
    for (int i = 0; i < 28; ++i) {
        out[i] = info28[i] & 1u;
        const uint32_t fb = ((rem >> 11) ^ out[i]) & 1u;
        rem = (rem << 1) & 0xFFFu;
        if (fb) rem ^= 0x539u;          // g(x) without its x^12 term
    }
    for (int i = 0; i < 12; ++i) out[28 + i] = (uint8_t)((rem >> (11 - i)) & 1u);
    return out;
}

static void head(Word28 &w, unsigned t1, unsigned t2, unsigned dcc) {
    w[0] = (uint8_t)t1; w[1] = (uint8_t)t2; w[2] = (dcc >> 1) & 1u; w[3] = dcc & 1u;
}

Word28 overhead_word_1(unsigned dcc, unsigned sid, bool ep, bool auth, bool pci, unsigned nawc) {
    Word28 w{};
    head(w, 1, 1, dcc);
    expandbits(&w[4], 14, sid >> 1);
    w[18] = ep; w[19] = auth; w[20] = pci;
    expandbits(&w[21], 4, nawc);
    w[25] = 1; w[26] = 1; w[27] = 0;
    return w;
}

Word28 overhead_word_2(unsigned dcc, bool s, bool e, bool regh, bool regr, unsigned dtx, unsigned nminusone, bool rcf,
                       bool cpa, unsigned cmax, bool end) {
    Word28 w{};
    head(w, 1, 1, dcc);
    w[4] = s; w[5] = e; w[6] = regh; w[7] = regr;
    w[8] = (dtx >> 1) & 1u; w[9] = dtx & 1u;
    expandbits(&w[10], 5, nminusone);
    w[15] = rcf; w[16] = cpa;
    expandbits(&w[17], 7, cmax);
    w[24] = end; w[25] = 1; w[26] = 1; w[27] = 1;
    return w;
}

Word28 control_filler_word() {
    static const char *b = "1100010111000001100111111001";     // lib/focc_impl.cc:293-295
    Word28 w{};
    for (int i = 0; i < 28; ++i) w[i] = (uint8_t)(b[i] - '0');
    return w;
}

Word28 access_type_parameters_global_action(unsigned dcc, bool end) {
    Word28 w{};
    head(w, 1, 1, dcc);
    w[4] = 1; w[7] = 1;                 // ACT = 1001
    w[24] = end; w[25] = 1;             // OHD = 100
    return w;
}

Word28 registration_increment_global_action(unsigned dcc, unsigned regincr, bool end) {
    Word28 w{};
    head(w, 1, 1, dcc);
    w[6] = 1;                           // ACT = 0010
    expandbits(&w[8], 12, regincr);
    w[24] = end; w[25] = 1;
    return w;
}

Word28 registration_id(unsigned dcc, unsigned long regid, bool end) {
    Word28 w{};
    head(w, 1, 1, dcc);
    expandbits(&w[4], 20, regid);
    w[24] = end;
    return w;
}

Word28 focc_word1(bool multiword, unsigned dcc, uint64_t min1) {
    Word28 w{};
    head(w, 0, multiword ? 1 : 0, dcc);
    expandbits(&w[4], 24, min1);
    return w;
}

Word28 focc_word2_general(uint64_t min2, unsigned msg_type, unsigned ordq, unsigned order) {
    Word28 w{};
    w[0] = 1; w[1] = 0; w[2] = 1; w[3] = 1;
    expandbits(&w[4], 10, min2);
    expandbits(&w[15], 5, msg_type);
    expandbits(&w[20], 3, ordq);
    expandbits(&w[23], 5, order);
    return w;
}

Word28 fvc_word1_general(unsigned pscc, unsigned msg_type, unsigned ordq, unsigned order) {
    Word28 w{};
    w[0] = 1; w[1] = 0; w[2] = 1; w[3] = 1;
    w[4] = (pscc >> 1) & 1u; w[5] = pscc & 1u;
    expandbits(&w[15], 5, msg_type);
    expandbits(&w[20], 3, ordq);
    expandbits(&w[23], 5, order);
    return w;
}

Word28 focc_word2_voice_channel(unsigned scc, uint64_t min2, unsigned vmac, unsigned chan) {
    Word28 w{};
    w[0] = 1; w[1] = 0; w[2] = (scc >> 1) & 1u; w[3] = scc & 1u;
    expandbits(&w[4], 10, min2);
    w[14] = (vmac >> 2) & 1u; w[15] = (vmac >> 1) & 1u; w[16] = vmac & 1u;
    expandbits(&w[17], 11, chan);
    return w;
}

uint64_t compute_min_3(char d1c, char d2c, char d3c) {
    uint64_t d[3] = {(uint64_t)(d1c - '0'), (uint64_t)(d2c - '0'), (uint64_t)(d3c - '0')};
    for (auto &x : d)
        if (x == 0) x = 10;                                   // digit 0 counts as ten (TIA-553 2.3.1)
    return 100 * d[0] + 10 * d[1] + d[2] - 111;
}

bool parse_min(const std::string &min, uint64_t &min1, uint64_t &min2) {
    if (min.size() != 10) return false;
    for (char c : min)
        if (c < '0' || c > '9') return false;
    min2 = compute_min_3(min[0], min[1], min[2]);
    uint64_t thousands = (uint64_t)(min[6] - '0');
    if (thousands == 0) thousands = 10;
    min1 = ((compute_min_3(min[3], min[4], min[5]) & 0x3ff) << 14) | ((thousands & 0xf) << 10) |
           (compute_min_3(min[7], min[8], min[9]) & 0x3ff);
    return true;
}

std::array<uint8_t, 463> focc_frame_slots(const uint8_t *wa, const uint8_t *wb) {
    static const uint8_t dot[10] = {1, 0, 1, 0, 1, 0, 1, 0, 1, 0};
    static const uint8_t sync[11] = {1, 1, 1, 0, 0, 0, 1, 0, 0, 1, 0};
    const auto ea = bch_encode_40_28(wa), eb = bch_encode_40_28(wb);
    std::array<uint8_t, 463> f{};
    int p = 0;
    f[p++] = kSlotBI; std::memcpy(&f[p], dot, 10); p += 10;
    f[p++] = kSlotBI; std::memcpy(&f[p], sync, 11); p += 11;
    for (int rep = 0; rep < 5; ++rep)
        for (int w = 0; w < 2; ++w) {
            const auto &e = w ? eb : ea;
            for (int q = 0; q < 4; ++q) {
                f[p++] = kSlotBI;
                std::memcpy(&f[p], &e[10 * q], 10);
                p += 10;
            }
        }
    return f;
}

std::vector<uint8_t> fvc_word_train(const uint8_t *word28) {
    static const uint8_t sync[11] = {1, 1, 1, 0, 0, 0, 1, 0, 0, 1, 0};
    const auto enc = bch_encode_40_28(word28);
    std::vector<uint8_t> t;
    t.reserve(1032);
    for (int i = 0; i < 101; ++i) t.push_back((uint8_t)((i & 1) ^ 1));
    for (int j = 0; j < 11; ++j) {
        t.insert(t.end(), sync, sync + 11);
        t.insert(t.end(), enc.begin(), enc.end());
        if (j < 10) for (int i = 0; i < 37; ++i) t.push_back((uint8_t)((i & 1) ^ 1));
    }
    return t;
}

}  // namespace amps
