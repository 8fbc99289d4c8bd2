/*
 * amps_words.c -- ORACLE (test infrastructure).  28-bit FOCC/FVC word builders and MIN arithmetic,
 * one byte per bit, MSB first.  Restates lib/focc_impl.cc:252-381, lib/amps_packet.cc:26-95,
 * lib/amps_packet.h:277-366, lib/utils.cc:101-108.
 */
#include "amps_oracle.h"
#include <string.h>

void orc_expandbits(uint8_t *out, size_t nbits, uint64_t val) {
    for (size_t i = 0; i < nbits; i++) out[i] = (uint8_t)((val >> (nbits - 1 - i)) & 1u);
}

static void t1t2_dcc(uint8_t w[28], unsigned t1, unsigned t2, unsigned dcc) {
    w[0] = (uint8_t)t1; w[1] = (uint8_t)t2;
    w[2] = (dcc >> 1) & 1u; w[3] = dcc & 1u;
}

/* lib/focc_impl.cc:252-269 */
void orc_overhead_word_1(uint8_t w[28], unsigned dcc, unsigned sid, int ep, int auth, int pci, unsigned nawc) {
    t1t2_dcc(w, 1, 1, dcc);
    orc_expandbits(w + 4, 14, sid >> 1);        /* SID1 = upper 14 bits of the 15-bit SID */
    w[18] = ep ? 1 : 0; w[19] = auth ? 1 : 0; w[20] = pci ? 1 : 0;
    orc_expandbits(w + 21, 4, nawc);
    w[25] = 1; w[26] = 1; w[27] = 0;            /* OHD = 110 */
}
/* lib/focc_impl.cc:270-292 */
void orc_overhead_word_2(uint8_t w[28], unsigned dcc, int s, int e, int regh, int regr, unsigned dtx,
                         unsigned nminusone, int rcf, int cpa, unsigned cmax, int end) {
    t1t2_dcc(w, 1, 1, dcc);
    w[4] = s ? 1 : 0; w[5] = e ? 1 : 0; w[6] = regh ? 1 : 0; w[7] = regr ? 1 : 0;
    w[8] = (dtx >> 1) & 1u; w[9] = dtx & 1u;
    orc_expandbits(w + 10, 5, nminusone);
    w[15] = rcf ? 1 : 0; w[16] = cpa ? 1 : 0;
    orc_expandbits(w + 17, 7, cmax);
    w[24] = end ? 1 : 0;
    w[25] = 1; w[26] = 1; w[27] = 1;            /* OHD = 111 */
}
/* lib/focc_impl.cc:293-295 */
void orc_control_filler_word(uint8_t w[28]) {
    static const char *bits = "1100010111000001100111111001";
    for (int i = 0; i < 28; i++) w[i] = (uint8_t)(bits[i] - '0');
}
/* lib/focc_impl.cc:296-335 */
void orc_access_type_global_action(uint8_t w[28], unsigned dcc, int end) {
    memset(w, 0, 28);
    t1t2_dcc(w, 1, 1, dcc);
    w[4] = 1; w[7] = 1;                         /* ACT = 1001 */
    w[24] = end ? 1 : 0;
    w[25] = 1;                                  /* OHD = 100 */
}
/* lib/focc_impl.cc:336-362 */
void orc_reg_increment_global_action(uint8_t w[28], unsigned dcc, unsigned regincr, int end) {
    memset(w, 0, 28);
    t1t2_dcc(w, 1, 1, dcc);
    w[6] = 1;                                   /* ACT = 0010 */
    orc_expandbits(w + 8, 12, regincr);
    w[24] = end ? 1 : 0;
    w[25] = 1;
}
/* lib/focc_impl.cc:365-381 */
void orc_registration_id(uint8_t w[28], unsigned dcc, unsigned long regid, int end) {
    memset(w, 0, 28);
    t1t2_dcc(w, 1, 1, dcc);
    orc_expandbits(w + 4, 20, regid);
    w[24] = end ? 1 : 0;                        /* OHD = 000 */
}
/* lib/amps_packet.cc:26-32 */
void orc_focc_word1(uint8_t w[28], int multiword, unsigned dcc, uint64_t min1) {
    t1t2_dcc(w, 0, multiword ? 1 : 0, dcc);
    orc_expandbits(w + 4, 24, min1);
}
/* lib/amps_packet.cc:38-49 */
void orc_focc_word2_general(uint8_t w[28], uint64_t min2, unsigned msg_type, unsigned ordq, unsigned order) {
    w[0] = 1; w[1] = 0; w[2] = 1; w[3] = 1;     /* T1T2 = 10, SCC = 11 */
    orc_expandbits(w + 4, 10, min2);
    w[14] = 0;
    orc_expandbits(w + 15, 5, msg_type);
    orc_expandbits(w + 20, 3, ordq);
    orc_expandbits(w + 23, 5, order);
}
/* lib/amps_packet.cc:55-76 */
void orc_fvc_word1_general(uint8_t w[28], unsigned pscc, unsigned msg_type, unsigned ordq, unsigned order) {
    memset(w, 0, 28);
    w[0] = 1; w[1] = 0; w[2] = 1; w[3] = 1;
    w[4] = (pscc >> 1) & 1u; w[5] = pscc & 1u;
    orc_expandbits(w + 15, 5, msg_type);
    orc_expandbits(w + 20, 3, ordq);
    orc_expandbits(w + 23, 5, order);
}
/* lib/amps_packet.cc:82-95 */
void orc_focc_word2_voice_channel(uint8_t w[28], unsigned scc, uint64_t min2, unsigned vmac, unsigned chan) {
    w[0] = 1; w[1] = 0; w[2] = (scc >> 1) & 1u; w[3] = scc & 1u;
    orc_expandbits(w + 4, 10, min2);
    w[14] = (vmac >> 2) & 1u; w[15] = (vmac >> 1) & 1u; w[16] = vmac & 1u;
    orc_expandbits(w + 17, 11, chan);
}

/* lib/amps_packet.h:277-302 -- three MIN digits from a 10-bit group (553 2.3.1) */
void orc_extract_min_3(uint64_t val, char out3[3]) {
    uint64_t m2 = val + 111;
    uint64_t dig = m2 % 10;
    out3[2] = (char)('0' + dig);
    m2 -= (dig == 0) ? 10 : dig;
    dig = (m2 % 100) / 10;
    out3[1] = (char)('0' + dig);
    if (dig == 0) m2 -= 100; else m2 -= (m2 % 100);
    dig = m2 / 100;
    if (dig > 9) dig = 0;
    out3[0] = (char)('0' + dig);
}
/* lib/amps_packet.h:305-319 */
uint64_t orc_compute_min_3(char d1c, char d2c, char d3c) {
    uint64_t d1 = (uint64_t)(d1c - '0'), d2 = (uint64_t)(d2c - '0'), d3 = (uint64_t)(d3c - '0');
    if (d1 == 0) d1 = 10;
    if (d2 == 0) d2 = 10;
    if (d3 == 0) d3 = 10;
    return 100 * d1 + 10 * d2 + d3 - 111;
}
/* lib/amps_packet.h:328-349 (the reference reads min[0..9] whatever the length; we require 10 digits
 * to stay in bounds and return 0 for the lengths on which the reference would read past the string) */
int orc_parse_min(const char *min, uint64_t *min1, uint64_t *min2) {
    size_t len = strlen(min);
    if (len < 1 || len > 10) return 0;
    for (size_t i = 0; i < len; i++) if (min[i] < '0' || min[i] > '9') return 0;
    if (len != 10) return 0;
    *min2 = orc_compute_min_3(min[0], min[1], min[2]);
    uint64_t thous = (uint64_t)(min[6] - '0');
    if (thous == 0) thous = 10;
    *min1 = ((orc_compute_min_3(min[3], min[4], min[5]) & 0x3ff) << 14) | ((thous & 0xf) << 10) |
            (orc_compute_min_3(min[7], min[8], min[9]) & 0x3ff);
    return 1;
}
/* lib/amps_packet.h:354-363 */
void orc_calc_min(uint64_t min1, uint64_t min2, char out11[11]) {
    orc_extract_min_3(min2, out11);
    orc_extract_min_3((min1 >> 14) & 0x3ff, out11 + 3);
    uint64_t thous = (min1 >> 10) & 0xf;
    if (thous > 9) thous = 0;
    out11[6] = (char)('0' + thous);
    orc_extract_min_3(min1 & 0x3ff, out11 + 7);
    out11[10] = 0;
}
