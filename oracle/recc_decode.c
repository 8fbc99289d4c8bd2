/*
 * recc_decode.c -- ORACLE (test infrastructure).  RECC burst (3374 hard half-symbols) -> DCC,
 * 7 x 240 word bits, Manchester error counts, BCH validity, parsed fields, message class.
 * Restates lib/utils.cc:27-59 (pair table), lib/recc_decode_impl.cc:81-169 (burst handler),
 * lib/amps_packet.h:103-274 (word parsers), :277-366 (MIN).
 *
 * Kept quirks: fields are parsed from the RAW first repeat of each word, not from the
 * BCH-corrected bits (recc_decode_impl.cc:112,117,130,146,161); only Word A's validity gates
 * (:108-111); E == 0 drops the message (:113-116).
 */
#include "amps_oracle.h"
#include <string.h>

size_t orc_manchester_decode(const uint8_t *src, uint8_t *dst, size_t dstlen) {
    size_t bad = 0;
    for (size_t o = 0; o < dstlen; o++) {
        unsigned a = src[2 * o] & 1u, b = src[2 * o + 1] & 1u;
        /* (1,0)->0  (0,1)->1  (1,1)->0+err  (0,0)->1+err */
        dst[o] = (uint8_t)(a ? 0 : 1);
        if (a == b) bad++;
    }
    return bad;
}

static uint64_t getbits(const uint8_t *b, int n) {
    uint64_t v = 0;
    for (int i = 0; i < n; i++) v = (v << 1) | (b[i] & 1u);
    return v;
}

/* lib/amps_packet.h:207-273 */
static void called_digits(uint32_t digits, char *out /* >= 9 */) {
    int n = 0;
    for (int i = 0; i < 8; i++) {
        unsigned v = (digits >> 28) & 0xf;
        if (v == 0 || v >= 13) break;
        out[n++] = (v <= 9) ? (char)('0' + v) : (v == 10 ? '0' : (v == 11 ? '*' : '#'));
        digits <<= 4;
    }
    out[n] = 0;
}

void orc_recc_decode(const uint8_t blob[3374], orc_recc_result *r) {
    memset(r, 0, sizeof *r);
    r->dcc_errs = (uint8_t)orc_manchester_decode(blob, r->dcc, 7);
    for (int w = 0; w < 7; w++)
        r->errs[w] = (uint16_t)orc_manchester_decode(blob + 14 + 480 * w, r->words[w], 240);
    for (int w = 0; w < 7; w++) {
        r->valid[w] = 0; r->valid_repeat[w] = 5;
        for (int rep = 0; rep < 5; rep++) {
            if (orc_bch_decode_48(r->words[w] + 48 * rep, NULL)) { r->valid[w] = 1; r->valid_repeat[w] = (uint8_t)rep; break; }
        }
    }
    const uint8_t *a = r->words[0], *b = r->words[1];
    r->F = a[0] & 1u; r->NAWC = (uint8_t)getbits(a + 1, 3);
    r->T = a[4] & 1u; r->S = a[5] & 1u; r->E = a[6] & 1u; r->ER = a[7] & 1u;
    r->SCM = (uint8_t)getbits(a + 8, 4);
    r->MIN1 = (uint32_t)getbits(a + 12, 24);
    r->B_F = b[0] & 1u; r->B_NAWC = (uint8_t)getbits(b + 1, 3);
    r->MSG_TYPE = (uint8_t)getbits(b + 4, 5); r->ORDQ = (uint8_t)getbits(b + 9, 3);
    r->ORDER = (uint8_t)getbits(b + 12, 5);
    r->LT = b[17] & 1u; r->EP = b[18] & 1u; r->SCM4 = b[19];
    r->MPCI = (uint8_t)getbits(b + 20, 2); r->SDCC1 = (uint8_t)getbits(b + 22, 2); r->SDCC2 = (uint8_t)getbits(b + 24, 2);
    r->MIN2 = (uint16_t)getbits(b + 26, 10);
    r->word_c_serial = (uint32_t)getbits(r->words[2] + 4, 32);
    orc_calc_min(r->MIN1, r->MIN2, r->min);

    if (!r->valid[0]) { r->kind = 0; return; }
    if (!r->E) { r->kind = 1; return; }
    int order_zero = (r->ORDER == 0 && r->ORDQ == 0 && r->MSG_TYPE == 0);
    if (r->T == 0 && order_zero) {
        r->kind = 2;                                    /* page response (:121) */
    } else if (r->T == 1 && r->ORDER == 0xd) {
        r->kind = 3;                                    /* registration (:123-138) */
        if (r->S && r->NAWC > 1) r->esn = r->word_c_serial;
    } else if (r->T == 1 && (r->NAWC > 2 || order_zero)) {
        unsigned nawc = r->NAWC;                        /* origination (:139-165) */
        unsigned next = 2;
        if (r->S) { r->esn = r->word_c_serial; next++; nawc = (unsigned)(uint8_t)(r->NAWC - 2); }
        if (nawc < 1 || nawc > 4) { r->kind = 6; return; }
        r->kind = 4;
        size_t len = 0;
        for (; nawc > 0; nawc--) {
            char d[9];
            called_digits((uint32_t)getbits(r->words[next] + 4, 32), d);
            next++;
            size_t dl = strlen(d);
            memcpy(r->dialed + len, d, dl); len += dl;
        }
        r->dialed[len] = 0;
    } else {
        r->kind = 5;
    }
}

/* What recc_decode_impl publishes for a decoded burst: restates handle_registration (:181-190),
 * handle_response (:195-220) and handle_origination (:234-272) of lib/recc_decode_impl.cc.  stream is always
 * STREAM_BOTH (the reference overrides its own A/B choice, :247). */
void orc_recc_actions_for(const orc_recc_result *r, orc_recc_actions *a) {
    memset(a, 0, sizeof *a);
    a->fvc_mute = -1; a->audio_mute = -1;
    if (r->kind == 3) {
        a->n_focc = 2; a->focc_stream = 3;
        orc_focc_word1(a->focc_words[0], 1, 0, r->MIN1);
        orc_focc_word2_general(a->focc_words[1], r->MIN2, 0, 0, 7);
    } else if (r->kind == 2) {
        a->n_focc = 2; a->focc_stream = 3;
        orc_focc_word1(a->focc_words[0], 1, 0, r->MIN1);
        orc_focc_word2_voice_channel(a->focc_words[1], 1, r->MIN2, 0, 355);
        a->has_fvc = 1; a->fvc_timer = 35;
        orc_fvc_word1_general(a->fvc_word, 1, 0, 0, 1);
        a->fvc_mute = 0; a->audio_mute = 1;
    } else if (r->kind == 4) {
        a->n_focc = 2; a->focc_stream = 3;
        orc_focc_word1(a->focc_words[0], 1, 0, r->MIN1);
        if (r->dialed[0] == '0') orc_focc_word2_general(a->focc_words[1], r->MIN2, 0, 0, 9);
        else orc_focc_word2_voice_channel(a->focc_words[1], 1, r->MIN2, 0, 356);
        a->fvc_mute = 1; a->audio_mute = 0;
        strcpy(a->command, "page ");
        strncat(a->command, r->dialed, sizeof a->command - 6);
    }
}
