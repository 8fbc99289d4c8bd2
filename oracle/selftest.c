/*
 * selftest.c -- ORACLE (test infrastructure).  A plain C driver that pushes every oracle routine through a small seeded
 * scenario; built with -fsanitize=address,undefined by `make -C oracle selftest_asan` and run by tests/test_oracle_cpu.py:
 * the checker itself must be free of out-of-bounds accesses, overflows and misaligned loads before it is trusted.
 * Exit status 0 = every internal consistency check held (the sanitizers abort on their own findings).
 */
#include "amps_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static uint64_t lcg = 0xA3B5;
static uint32_t rnd(void) { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(lcg >> 33); }
static int fails = 0;
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "selftest: %s failed (line %d)\n", #c, __LINE__); fails++; } } while (0)

static int nblobs = 0;
static uint8_t last_blob[ORC_RECC_CAPTURE_LEN];
static void on_blob(const uint8_t *b, void *u) { (void)u; memcpy(last_blob, b, ORC_RECC_CAPTURE_LEN); nblobs++; }

/* a RECC origination message as hard half-symbols: 30-bit dotting, word sync, DCC, 7 words x 5 repeats, Manchester */
static size_t make_halfsyms(uint8_t *hs) {
    static const uint8_t sync[11] = {1,1,1,0,0,0,1,0,0,1,0};
    uint8_t bits[48 + 7 * 240];
    size_t nb = 0;
    for (int i = 0; i < 30; i++) bits[nb++] = (uint8_t)((i & 1) ^ 1);
    for (int i = 0; i < 11; i++) bits[nb++] = sync[i];
    for (int i = 0; i < 7; i++) bits[nb++] = 0;
    for (int w = 0; w < 7; w++) {
        uint8_t info[36], enc[48];
        for (int i = 0; i < 36; i++) info[i] = (uint8_t)(rnd() & 1);
        if (w == 0) { info[0] = 1; info[1] = 1; info[2] = 1; info[3] = 0; info[4] = 1; info[5] = 1; info[6] = 1; info[7] = 0; }   /* F=1 NAWC=6 T=1 S=1 E=1 */
        orc_bch_encode_48_36(info, enc);
        CHECK(orc_bch_decode_48(enc, NULL) == 1);
        for (int r = 0; r < 5; r++) for (int i = 0; i < 48; i++) bits[nb++] = enc[i];
    }
    for (size_t i = 0; i < nb; i++) { hs[2 * i] = (uint8_t)(1 - bits[i]); hs[2 * i + 1] = bits[i]; }
    return 2 * nb;
}

int main(void) {
    /* ---- BCH + words */
    uint8_t w[28], c40[40];
    orc_overhead_word_1(w, 0, 16, 1, 0, 0, 3); orc_bch_encode_40_28(w, c40);
    orc_overhead_word_2(w, 0, 1, 1, 1, 1, 0, 23, 1, 1, 23, 0); orc_bch_encode_40_28(w, c40);
    orc_control_filler_word(w); orc_access_type_global_action(w, 0, 0); orc_reg_increment_global_action(w, 0, 100, 0);
    orc_registration_id(w, 0, 500, 1); orc_focc_word1(w, 1, 0, 0xABCDE); orc_focc_word2_general(w, 0x155, 0, 0, 7);
    orc_fvc_word1_general(w, 1, 0, 0, 1); orc_focc_word2_voice_channel(w, 1, 0x2AA, 0, 355);
    uint64_t m1, m2; char minstr[11];
    CHECK(orc_parse_min("2125551234", &m1, &m2) == 1);
    orc_calc_min(m1, m2, minstr);
    CHECK(strncmp(minstr, "2125551234", 10) == 0);
    CHECK(orc_parse_min("123", &m1, &m2) == 0);
    for (int i = 0; i < 2000; i++) {                       /* random 48-bit words through the decoder */
        uint8_t r[48], o[48];
        for (int k = 0; k < 48; k++) r[k] = (uint8_t)(rnd() & 1);
        (void)orc_bch_decode_48(r, o);
    }
    /* ---- FOCC / FVC sources under ragged requests and injections */
    orc_focc *f = orc_focc_new(100000, 1);
    uint8_t *buf = (uint8_t *)malloc(70000);
    size_t total = 0;
    for (int i = 0; i < 400; i++) {
        if (i % 37 == 5) { uint8_t ww[56]; for (int k = 0; k < 56; k++) ww[k] = (uint8_t)(rnd() & 1); orc_focc_push_words(f, 1 + (long)(rnd() % 3), ww, 2); }
        int n = (int)(rnd() % 9000);
        int r = orc_focc_work(f, buf, n);
        CHECK(r <= n && (n < 1 ? r == -1 : r >= 0));
        if (r > 0) total += (size_t)r;
    }
    CHECK(total > 20000 && orc_focc_superframe_frames(f) >= 19);
    orc_focc_free(f);
    orc_fvc *v = orc_fvc_new(100000);
    int off = 0;
    memset(buf, 0x77, 64);
    CHECK(orc_fvc_work(v, buf, 64, &off) == 64 && buf[0] == 0x77);            /* idle: untouched */
    orc_fvc_word1_general(w, 1, 0, 0, 1);
    orc_fvc_push_words(v, w, 1, 1, 3);
    for (int i = 0; i < 60; i++) { int r = orc_fvc_work(v, buf, 1 + (int)(rnd() % 20000), &off); CHECK(r > 0); }
    orc_fvc_free(v);
    /* ---- RECC capture under random chunkings + decode + responses */
    uint8_t *hs = (uint8_t *)malloc(8000), *stream = (uint8_t *)malloc(200000);
    size_t nhs = make_halfsyms(hs), ns = 0;
    for (int rep = 0; rep < 12; rep++) {
        size_t gap = 100 + rnd() % 9000;
        for (size_t i = 0; i < gap; i++) stream[ns++] = (uint8_t)(rnd() & 1);
        memcpy(stream + ns, hs, nhs); ns += nhs;
    }
    for (size_t i = 0; i < 5000; i++) stream[ns++] = (uint8_t)(rnd() & 1);
    orc_recc *rc = orc_recc_new();
    for (size_t pos = 0; pos < ns;) {
        size_t n = 1 + rnd() % 61000;
        if (n > ns - pos) n = ns - pos;
        CHECK(orc_recc_work(rc, stream + pos, (int)n, on_blob, NULL) == 0);
        pos += n;
    }
    CHECK(orc_recc_work(rc, stream, 61440, on_blob, NULL) == -2);
    CHECK(nblobs >= 1 && orc_recc_buflen(rc) <= 65536);
    orc_recc_free(rc);
    orc_recc_result res;
    orc_recc_actions act;
    orc_recc_decode(last_blob, &res);
    orc_recc_actions_for(&res, &act);
    for (int i = 0; i < 50; i++) {                           /* garbage blobs through the decoder and the dispatcher */
        uint8_t blob[ORC_RECC_CAPTURE_LEN];
        for (int k = 0; k < ORC_RECC_CAPTURE_LEN; k++) blob[k] = (uint8_t)(rnd() & 1);
        orc_recc_decode(blob, &res);
        orc_recc_actions_for(&res, &act);
    }
    orc_cmd_actions ca;
    const char *cmds[] = {"page 2125551234", "page 12", "page ", "fvc on", "fvc off", "fvc alert", "", "PAGE  9075550199\n", "bogus"};
    for (size_t i = 0; i < sizeof cmds / sizeof cmds[0]; i++) orc_command_actions(cmds[i], &ca);
    /* ---- the float chains on a short FM burst at 10 MS/s and 400 kS/s */
    float taps[299];
    CHECK(orc_firdes_low_pass(3.0, 400e3, 10e3, 4500.0, 2, taps, 299) == 299);
    const size_t n = 50 * 9000;                              /* 450 000 samples: trigger + part of the message */
    float *iq = (float *)malloc(sizeof(float) * 2 * n);
    double phi = 0;
    for (size_t i = 0; i < n; i++) {
        size_t k = i / 500;
        double s = (i < 20000 || k - 40 >= nhs) ? 0.0 : (hs[k - 40] ? 1.0 : -1.0);
        phi += s * 2.0 * M_PI * 8000.0 / 10e6;
        double ang = 2.0 * M_PI * fmod(-0.016 * (double)i, 1.0) + phi;
        double nr = ((double)rnd() / 2147483648.0 - 0.5) * 0.2, ni = ((double)rnd() / 2147483648.0 - 0.5) * 0.2;
        iq[2 * i] = (float)((s != 0.0 ? 0.5 * cos(ang) : 0.0) + nr);
        iq[2 * i + 1] = (float)((s != 0.0 ? 0.5 * sin(ang) : 0.0) + ni);
    }
    const uint32_t fcw = orc_nco_fcw(-160e3, 10e6);
    float *y32 = (float *)malloc(sizeof(float) * 2 * (n / 50)), *d32 = (float *)malloc(sizeof(float) * (n / 50));
    double *y64 = (double *)malloc(sizeof(double) * 2 * (n / 50)), *d64 = (double *)malloc(sizeof(double) * (n / 50));
    orc_rx_chain_f32(iq, n, fcw, taps, 299, y32, d32);
    orc_rx_chain_f64(iq, n, fcw, taps, 299, y64, d64);
    double err = 0;
    for (size_t i = 0; i < 2 * (n / 50); i++) err += (y32[i] - y64[i]) * (y32[i] - y64[i]);
    CHECK(sqrt(err / (double)(n / 50)) < 1e-6);
    orc_rx_chain_f32_at(iq, n, fcw, taps, 299, 123456789u, y32, d32);
    orc_rx_chain400_f32(iq, 2 * 9000, orc_nco_fcw(-160e3, 400e3), taps, 299, y32, d32);
    orc_rx_chain400_f64(iq, 2 * 9000, orc_nco_fcw(-160e3, 400e3), taps, 299, y64, d64);
    orc_quad_demod(iq, 1000, d32, d64);
    orc_rx_chain_f32(iq, n, fcw, taps, 299, NULL, d32);
    orc_burst *bursts = (orc_burst *)malloc(sizeof(orc_burst) * 4);
    CHECK(orc_rx_detect(d32, n / 50, bursts, 4) == 0);        /* the capture is not complete in 9000 demod samples */
    CHECK(orc_rx_detect(d32, 100, bursts, 4) == 0);
    float *T = (float *)malloc(sizeof(float) * 129 * 8);
    orc_mmse_table(T);
    orc_mm_state ms;
    orc_mm_init(&ms);
    uint8_t *sym = (uint8_t *)malloc(n / 50 / 8 + 64);
    size_t nsym = orc_mm_process(&ms, d32, 3000, T, sym, n / 50 / 8 + 64);
    nsym += orc_mm_process(&ms, d32, n / 50, T, sym, n / 50 / 8 + 64);
    CHECK(nsym > 800 && nsym < 1000);
    /* ---- forward chain + voice leg */
    const size_t nsy = 1500;
    int8_t *s0 = (int8_t *)malloc(nsy), *s1 = (int8_t *)malloc(nsy);
    for (size_t i = 0; i < nsy; i++) { s0[i] = (i / 5) & 1 ? 1 : -1; s1[i] = i > 700 && i < 900 ? 0 : ((i / 5) % 3 ? 1 : -1); }
    float t0[193], t1[321];
    CHECK(orc_firdes_low_pass(1.0, 400e3, 10e3, 5e3, 0, t0, 193) == 193 && orc_firdes_low_pass(1.0, 400e3, 10e3, 3e3, 0, t1, 321) == 321);
    const int8_t *syms[2] = {s0, s1};
    const float *tp[2] = {t0, t1};
    const int nt[2] = {193, 321};
    const uint32_t mix[2] = {0u, orc_nco_fcw(-60e3, 10e6)};
    double *out = (double *)malloc(sizeof(double) * 2 * nsy * 100);
    orc_fwd_chain_f64(syms, 2, nsy, (uint32_t)llround(8000.0 / 100e3 * 4294967296.0), tp, nt, mix, 0.5, out);
    float vt[225];
    CHECK(orc_firdes_low_pass(3.0, 400e3, 15e3, 6e3, 2, vt, 225) == 225);
    const size_t na = 4 * nsy / 25;
    float *audio = (float *)malloc(sizeof(float) * na);
    uint8_t *mute = (uint8_t *)calloc(na, 1);
    for (size_t i = 0; i < na; i++) { audio[i] = (float)(0.2 * sin(2.0 * M_PI * 440.0 * (double)i / 16000.0)); mute[i] = i > na / 2; }
    double *v400 = (double *)malloc(sizeof(double) * 2 * 25 * na);
    orc_voice_tx_f64(audio, na, 0.05, mute, vt, 225, v400);
    const double *extra[2] = {NULL, v400};
    orc_fwd_chain_voice_f64(syms, 2, nsy, (uint32_t)llround(8000.0 / 100e3 * 4294967296.0), tp, nt, mix, 0.5, extra, out);
    double pa[2], pb[2], E[25 * 29];
    orc_fm_preemph_taps(16000.0, 75e-6, -1.0, pb, pa);
    CHECK(orc_arb25_taps(vt, 225, E) > 0);
    free(buf); free(hs); free(stream); free(iq); free(y32); free(d32); free(y64); free(d64); free(bursts); free(T); free(sym);
    free(s0); free(s1); free(out); free(audio); free(mute); free(v400);
    printf("selftest: %d failed checks, %d blobs\n", fails, nblobs);
    return fails ? 1 : 0;
}
