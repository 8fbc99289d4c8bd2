/*
 * amps_oracle.h -- CPU ORACLE for the gr-amps hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This directory is a plain-C restatement of the reference's algorithm for the
 * path named by BASELINE.json:north_star.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (gr_amps_b200/) never links, imports or calls anything in here.
 *
 * PARITY STATUS.  Integer paths (BCH, words, FOCC/FVC sources, RECC capture, RECC decode + responses, command
 * processor): PINNED to the reference's own code -- oracle/_ref/libamps_ref.so is gr-amps's unmodified lib/ sources compiled
 * from /root/reference against stand-in GNU Radio / Boost / IT++ headers (oracle/Makefile, oracle/ref_harness.cc), and
 * tests/test_ref_pin_cpu.py requires byte-identical transcripts, live and through tests/golden/ref_vectors.json.
 * Floating paths (dsp_chain.c, mm_timing.c, voice_tx.c) restate stock GNU Radio 3.7 blocks whose source is not in the
 * reference tree, from their documented equations (grc/ampsbs.grc gives the parameters); the reference ships no golden
 * vectors and an empty test-suite (reference lib/qa_amps.cc:9-15).  They are held against (a) an independent numpy/scipy
 * implementation of the same blocks in GNU Radio's own structure (tests/scipy_chain.py, tests/test_independent_chain_*.py)
 * and (b) GNU Radio's own QA known answers for firdes.low_pass, quadrature_demod_cf and the M&M interpolator's DC gain
 * (tests/golden/kat_gnuradio_*.json).  Still "parity unpinned": GNU Radio's fast_atan2f table, the rest of its MMSE table,
 * VOLK's summation order, fm_preemph / pfb.arb_resampler of the voice leg, and IT++'s BCH decoder (restated, not linked).
 * The checker itself runs clean under -fsanitize=address,undefined (selftest.c, `make selftest_asan`).
 *
 * All citations are relative to /root/reference/.
 */
#ifndef AMPS_ORACLE_H
#define AMPS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- BCH (bch63.c) */
/* itpp::BCH(63, 2, true): g(x) = 0o12471 (lib/focc_impl.cc:105, lib/recc_decode_impl.cc:33) */
void     orc_bch_encode_40_28(const uint8_t info28[28], uint8_t out40[40]);   /* lib/focc_impl.cc:156-176 */
void     orc_bch_encode_48_36(const uint8_t info36[36], uint8_t out48[48]);   /* mobile side; same code, 15 pad zeros */
/* lib/recc_decode_impl.cc:53-79: pad 15 zeros, itpp BCH::decode; returns validity (1/0) and
 * writes the (possibly corrected) 48-bit word to out48 (may be NULL). */
int      orc_bch_decode_48(const uint8_t in48[48], uint8_t out48[48]);
uint32_t orc_bch_syndromes63(const uint8_t r63[63]);   /* (s1 | s3<<8) over GF(64), alpha from x^6+x+1 */

/* ---------------------------------------------------------------- words (amps_words.c) */
void orc_expandbits(uint8_t *out, size_t nbits, uint64_t val);                /* lib/utils.cc:101-108 */
void orc_overhead_word_1(uint8_t w[28], unsigned dcc, unsigned sid, int ep, int auth, int pci, unsigned nawc);
void orc_overhead_word_2(uint8_t w[28], unsigned dcc, int s, int e, int regh, int regr, unsigned dtx,
                         unsigned nminusone, int rcf, int cpa, unsigned cmax, int end);
void orc_control_filler_word(uint8_t w[28]);
void orc_access_type_global_action(uint8_t w[28], unsigned dcc, int end);
void orc_reg_increment_global_action(uint8_t w[28], unsigned dcc, unsigned regincr, int end);
void orc_registration_id(uint8_t w[28], unsigned dcc, unsigned long regid, int end);
void orc_focc_word1(uint8_t w[28], int multiword, unsigned dcc, uint64_t min1);               /* lib/amps_packet.cc:26-32 */
void orc_focc_word2_general(uint8_t w[28], uint64_t min2, unsigned msg_type, unsigned ordq, unsigned order);
void orc_fvc_word1_general(uint8_t w[28], unsigned pscc, unsigned msg_type, unsigned ordq, unsigned order);
void orc_focc_word2_voice_channel(uint8_t w[28], unsigned scc, uint64_t min2, unsigned vmac, unsigned chan);
/* MIN arithmetic, lib/amps_packet.h:277-366 */
void     orc_extract_min_3(uint64_t val, char out3[3]);
uint64_t orc_compute_min_3(char d1, char d2, char d3);
int      orc_parse_min(const char *min, uint64_t *min1, uint64_t *min2);
void     orc_calc_min(uint64_t min1, uint64_t min2, char out11[11]);

/* ---------------------------------------------------------------- FOCC source (focc_src.c) */
typedef struct orc_focc orc_focc;
orc_focc *orc_focc_new(unsigned long symrate, int aggressive_registration);   /* lib/focc_impl.cc:104-136 */
void      orc_focc_free(orc_focc *);
/* lib/focc_impl.cc:582-647: returns items produced (may be < n, incl. 0), -1 if n < 1 */
int       orc_focc_work(orc_focc *, uint8_t *out, int noutput_items);
/* lib/focc_impl.cc:521-563: stream 1=A 2=B 3=BOTH; words28 = nwords*28 bytes */
int       orc_focc_push_words(orc_focc *, long stream, const uint8_t *words28, long nwords);
int       orc_focc_superframe_frames(const orc_focc *);

/* ---------------------------------------------------------------- FVC source (fvc_src.c) */
typedef struct orc_fvc orc_fvc;
orc_fvc *orc_fvc_new(unsigned long symrate);                                  /* lib/fvc_impl.cc:56-67 */
void     orc_fvc_free(orc_fvc *);
/* lib/fvc_impl.cc:109-143; has_timer != 0 sets timerhack */
int      orc_fvc_push_words(orc_fvc *, const uint8_t *words28, long nwords, int has_timer, uint64_t timer);
/* lib/fvc_impl.cc:152-193; *fvc_off set to 1 when the "fvc off" PDU would be published.
 * When no word was ever pushed returns n and leaves out untouched (as the reference). */
int      orc_fvc_work(orc_fvc *, uint8_t *out, int noutput_items, int *fvc_off);

/* ---------------------------------------------------------------- RECC capture (recc_capture.c) */
#define ORC_RECC_TRIGGER_LEN 74
#define ORC_RECC_CAPTURE_LEN 3374
typedef struct orc_recc orc_recc;
typedef void (*orc_burst_cb)(const uint8_t *blob3374, void *user);
orc_recc *orc_recc_new(void);                                                 /* lib/recc_impl.cc:67-83 */
void      orc_recc_free(orc_recc *);
void      orc_recc_trigger(uint8_t out74[74]);                                /* lib/recc_impl.cc:51-65,76-79 */
/* lib/recc_impl.cc:93-145 (compat: buffer quirks reproduced). returns 0; -2 if n >= 61440 */
int       orc_recc_work(orc_recc *, const uint8_t *in, int n, orc_burst_cb cb, void *user);
size_t    orc_recc_buflen(const orc_recc *);

/* ---------------------------------------------------------------- RECC decode (recc_decode.c) */
typedef struct {
    uint8_t  dcc[7];
    uint8_t  dcc_errs;
    uint8_t  words[7][240];      /* Manchester-decoded bits, 5 repeats x 48 (lib/recc_decode_impl.cc:94-99) */
    uint16_t errs[7];            /* invalid Manchester pairs per word */
    uint8_t  valid[7];           /* BCH validity (first valid repeat) */
    uint8_t  valid_repeat[7];    /* index of the repeat that validated, 5 if none */
    /* word A (lib/amps_packet.h:154-161), parsed from RAW repeat 0 */
    uint8_t  F, NAWC, T, S, E, ER, SCM;
    uint32_t MIN1;
    /* word B (lib/amps_packet.h:177-188) */
    uint8_t  B_F, B_NAWC, MSG_TYPE, ORDQ, ORDER, LT, EP, SCM4, MPCI, SDCC1, SDCC2;
    uint16_t MIN2;
    uint32_t word_c_serial;      /* get32(words[2]+4,32) */
    /* dispatch result (lib/recc_decode_impl.cc:108-168) */
    int32_t  kind;               /* 0 none(invalid A), 1 E=0 drop, 2 page response, 3 registration, 4 origination, 5 unknown, 6 bad NAWC */
    uint32_t esn;
    char     min[11];
    char     dialed[33];
} orc_recc_result;
size_t orc_manchester_decode(const uint8_t *src, uint8_t *dst, size_t dstlen); /* lib/utils.cc:27-59 */
void   orc_recc_decode(const uint8_t blob[3374], orc_recc_result *out);        /* lib/recc_decode_impl.cc:81-169 */
typedef struct {
    int32_t  n_focc;             /* words in the focc_words tuple (0 = none published) */
    int64_t  focc_stream;
    uint8_t  focc_words[2][28];
    int32_t  has_fvc;            /* fvc_words tuple published */
    uint8_t  fvc_word[28];
    uint64_t fvc_timer;
    int32_t  fvc_mute, audio_mute;   /* -1 = not published, else the bool */
    char     command[48];        /* command_out PDU text, "" = none */
} orc_recc_actions;
void   orc_recc_actions_for(const orc_recc_result *r, orc_recc_actions *a);     /* lib/recc_decode_impl.cc:181-272 */
typedef struct {                 /* what amps.command_processor publishes for one command (command_proc.c) */
    int32_t  n_focc;
    int64_t  focc_stream;
    uint8_t  focc_words[2][28];
    int32_t  has_fvc;            /* fvc_words tuple (1, word), no timer */
    uint8_t  fvc_word[28];
    int32_t  fvc_mute, audio_mute;   /* -1 = not published */
    int32_t  n_debug;
    char     debug[2][48];       /* debug_output PDUs, in order */
} orc_cmd_actions;
void   orc_command_actions(const char *cmd, orc_cmd_actions *a);                /* lib/command_processor_impl.cc:52-117 */

/* ---------------------------------------------------------------- DSP chain (dsp_chain.c) */
/* gr::filter::firdes::low_pass (EXTERNAL). window: 0 hamming, 1 hann, 2 blackman.
 * returns ntaps; taps (fp32) written if taps != NULL and cap >= ntaps. */
int  orc_firdes_low_pass(double gain, double fs, double fc, double tw, int window, float *taps, int cap);

/* The 10 MS/s RX chain ("kernel-spec"):
 *   NCO(fcw, 32-bit phase) -> CIC^3 decimate-by-25 -> 299-tap firdes LPF decimate-by-2
 *   -> quadrature demod (atan2 of y[q]*conj(y[q-1])).
 * f64 flavour: straightforward double arithmetic, libm sin/cos/atan2 ("ideal").
 * f32 flavour: the exact fp32 op order of the CUDA kernel (bit-exact hard decisions). */
uint32_t orc_nco_fcw(double center_freq, double samp_rate);
/* iq: interleaved float re,im, n samples (n % 50 == 0). y (interleaved, n/50 complex) and d (n/50) */
void orc_rx_chain_f64(const float *iq, size_t n, uint32_t fcw, const float *h2, int nh2,
                      double *y_out, double *d_out);
void orc_rx_chain_f32(const float *iq, size_t n, uint32_t fcw, const float *h2, int nh2,
                      float *y_out, float *d_out);
void orc_quad_demod(const float *iq, size_t n, float *d32, double *d64);   /* quadrature_demod_cf alone: kernel-spec f32 and libm f64 */
void orc_rx_chain_f32_at(const float *iq, size_t n, uint32_t fcw, const float *h2, int nh2, uint32_t blk0,
                         float *y_out, float *d_out);   /* the same for a buffer that starts at absolute block blk0 of its stream */
/* native 400 kS/s front end (the reference's own rate): NCO -> lpf_taps /2 -> quadrature demod */
void orc_rx_chain400_f64(const float *iq, size_t n, uint32_t fcw, const float *h2, int nh2, double *y_out, double *d_out);
void orc_rx_chain400_f32(const float *iq, size_t n, uint32_t fcw, const float *h2, int nh2, float *y_out, float *d_out);
/* Trigger search + soft-peak timing + capture on the demod stream d (200 kS/s, 10 samples/half-symbol).
 * Fixed-mode semantics.  Writes up to max records; returns count. */
typedef struct {
    uint64_t d_index;            /* index into d of the first trigger half-symbol sample */
    float    corr;
    uint8_t  symbols[3374];
} orc_burst;
int  orc_rx_detect(const float *d, size_t nd, orc_burst *out, int max);

/* ---------------------------------------------------------------- serial M&M timing tail (mm_timing.c) */
typedef struct { float mu, omega, last; uint32_t pad; uint64_t pos; } orc_mm_state;
void   orc_mmse_table(float *T /* 129 x 8 */);
void   orc_mm_init(orc_mm_state *st);
size_t orc_mm_process(orc_mm_state *st, const float *d, uint64_t total, const float *T, uint8_t *sym, size_t cap);

/* Forward (TX) chain, f64 "ideal" (config 3); see dsp_chain.c for the definition. */
void orc_fwd_chain_f64(const int8_t *const *sym, int ncarriers, size_t nsym, uint32_t fcw_fm,
                       const float *const *taps, const int *ntaps, const uint32_t *fcw_mix, double scale,
                       double *out /* nsym*100 complex */);
void orc_fwd_chain_voice_f64(const int8_t *const *sym, int ncarriers, size_t nsym, uint32_t fcw_fm,
                             const float *const *taps, const int *ntaps, const uint32_t *fcw_mix, double scale,
                             const double *const *extra400, double *out);
/* ---------------------------------------------------------------- voice leg of the forward graph (voice_tx.c) */
void orc_fm_preemph_taps(double fs, double tau, double fh, double b[2], double a[2]);
int  orc_arb25_taps(const float *taps, int ntaps, double *E);
void orc_voice_tx_f64(const float *audio, size_t n_a, double sat_amp, const uint8_t *mute, const float *taps, int ntaps,
                      double *out /* 25 * n_a complex */);

#ifdef __cplusplus
}
#endif
#endif
