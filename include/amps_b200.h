/*
 * amps_b200.h -- C ABI of libamps_b200.so: the B200-native drop-in for gr-amps's per-sample DSP
 * hot path.  Plain pointers and sizes only; no C++/torch/GNU Radio types cross this boundary.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the gr-amps
 * tree).  A handle mirrors one GNU Radio block instance: like a block's work()/message handlers
 * (one scheduler thread per block), a handle must be used by one thread at a time.
 *
 * Every function returns an int status: AMPS_OK (0) or a negative AMPS_E_* code; nothing throws.
 * There is NO CPU fallback: without a usable sm_100 device every create() fails with
 * AMPS_E_NODEVICE and the library says so loudly on stderr.
 */
#ifndef AMPS_B200_H
#define AMPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AMPS_B200_API __attribute__((visibility("default")))
#else
#define AMPS_B200_API
#endif

#define AMPS_OK            0
#define AMPS_E_INVAL      (-1)   /* bad argument */
#define AMPS_E_NODEVICE   (-2)   /* no CUDA device / not sm_100 */
#define AMPS_E_CUDA       (-3)   /* CUDA runtime error (see amps_b200_last_error) */
#define AMPS_E_NOMEM      (-4)
#define AMPS_E_ALIGN      (-5)   /* device pointer / length alignment not met */
#define AMPS_E_OVERFLOW   (-6)   /* more samples than the handle was created for */
#define AMPS_E_STATE      (-7)

#define AMPS_RECC_TRIGGER_SYMS 74      /* lib/recc_impl.cc:76-77 */
#define AMPS_RECC_CAPTURE_SYMS 3374    /* lib/recc_impl.cc:70 */
#define AMPS_WORD_BITS 28              /* FOCC/FVC info word, one byte per bit (lib/amps_packet.cc) */

AMPS_B200_API int         amps_b200_version(void);
AMPS_B200_API const char *amps_b200_strerror(int status);
AMPS_B200_API const char *amps_b200_last_error(void);          /* thread-local detail string */
AMPS_B200_API int         amps_b200_device_count(void);        /* number of sm_100 devices, <0 on error */

/* ------------------------------------------------------------------------------------------
 * RECC word decode result -- replaces the locals of recc_decode_impl::bursts_message
 * (lib/recc_decode_impl.cc:81-169) and the parsers of lib/amps_packet.h:103-274.
 * Fields are parsed from the RAW first repeat of each word, exactly as the reference does.
 * ------------------------------------------------------------------------------------------ */
typedef struct amps_recc_words {
    uint8_t  dcc[7];
    uint8_t  dcc_errs;
    uint8_t  words[7][240];      /* Manchester-decoded bits, 5 repeats x 48 */
    uint16_t errs[7];            /* invalid Manchester pairs per word (lib/utils.cc:36-44) */
    uint8_t  valid[7];           /* BCH(63,51) validity of the first repeat that decodes */
    uint8_t  valid_repeat[7];    /* which repeat validated, 5 if none */
    uint8_t  F, NAWC, T, S, E, ER, SCM;
    uint8_t  pad0;
    uint32_t MIN1;
    uint8_t  B_F, B_NAWC, MSG_TYPE, ORDQ, ORDER, LT, EP, SCM4, MPCI, SDCC1, SDCC2;
    uint8_t  pad1;
    uint16_t MIN2;
    uint16_t pad2;
    uint32_t word_c_serial;
    int32_t  kind;               /* AMPS_MSG_* */
    uint32_t esn;
    char     min[12];            /* 10 digits + NUL */
    char     dialed[36];         /* up to 32 digits + NUL */
} amps_recc_words;

#define AMPS_MSG_INVALID_A     0   /* Word A failed BCH: dropped (recc_decode_impl.cc:108-111) */
#define AMPS_MSG_E0_DROPPED    1   /* E == 0 (:113-116) */
#define AMPS_MSG_PAGE_RESPONSE 2   /* :121 */
#define AMPS_MSG_REGISTRATION  3   /* :123-138 */
#define AMPS_MSG_ORIGINATION   4   /* :139-165 */
#define AMPS_MSG_UNKNOWN       5   /* :166-168 */
#define AMPS_MSG_BAD_NAWC      6   /* :154-157 */

/* One captured burst: what recc_impl::work publishes on message port "bursts"
 * (lib/recc_impl.cc:126: a 3374-byte blob of 0/1 half-symbols) plus where it was found and its
 * decode (what recc_decode_impl would compute from that blob). */
typedef struct amps_burst {
    uint64_t sample_index;       /* absolute input-sample index of the sampling instant of the first trigger half-symbol */
    uint64_t demod_index;        /* same, in demodulated samples (200 kS/s) */
    float    corr;               /* soft correlation of the 74-symbol trigger at the chosen phase */
    uint32_t run_length;         /* number of adjacent sampling phases that matched 74/74 */
    uint8_t  symbols[AMPS_RECC_CAPTURE_SYMS];
    uint8_t  pad[2];
    amps_recc_words decoded;
} amps_burst;

/* sizeof(amps_burst) / sizeof(amps_recc_words) as this library was built: lets a binding check its
 * struct layout before trusting it. */
AMPS_B200_API int amps_b200_abi_sizes(size_t *burst_bytes, size_t *words_bytes);

/* ------------------------------------------------------------------------------------------
 * Fused RECC receive path on IQ.  One handle replaces, for one carrier, the chain
 *   freq_xlating_fir_filter_ccc -> quadrature_demod_cf -> clock_recovery_mm_ff ->
 *   binary_slicer_fb (grc/ampsbs.grc:1814-1872, 774-816, 1751-1813, 1712-1750)
 *   -> amps.recc (lib/recc_impl.cc:93-145) -> burst decode (lib/recc_decode_impl.cc:81-169)
 * at samp_rate = 10 MS/s (25 x the reference's 400 kS/s) or at the reference's own
 * 400 kS/s (no extrapolation stage: NCO + lpf_taps /2 + quadrature demod are then exactly the reference's blocks).
 * ------------------------------------------------------------------------------------------ */
typedef struct amps_recc_iq amps_recc_iq;

typedef struct amps_recc_iq_params {
    double   samp_rate;          /* 10e6 (25 x the reference's rate; CIC^3 /25 first stage) or 400e3 (the reference's own rate, grc/ampsbs.grc:263) */
    double   center_freq;        /* carrier offset inside the band, Hz (reference: rx_offset = -160e3, grc/ampsbs.grc:212-238) */
    int      device;             /* CUDA ordinal */
    uint32_t max_samples;        /* largest nsamples ever passed in one call (sizes device buffers) */
    uint32_t max_bursts;         /* length of the pinned host ring of burst records; uncollected bursts beyond it are overwritten (0 -> 256) */
    uint32_t flags;              /* AMPS_RX_* */
    const float *lpf_taps;       /* channel filter taps @400 kS/s, NULL -> firdes.low_pass(3,400e3,10e3,4.5e3,BLACKMAN) */
    uint32_t n_lpf_taps;         /* <= 299 */
    float    sc16_scale;         /* AMPS_RX_INPUT_SC16: x = (float)int16 * sc16_scale per component; 0 -> 1/32768 */
} amps_recc_iq_params;

#define AMPS_RX_DUMP_BASEBAND 1u   /* keep the 200 kS/s complex baseband of the last call for inspection */
#define AMPS_RX_TIME_KERNELS  2u   /* bracket every front-end kernel launch with CUDA events (roofline accounting) */
#define AMPS_RX_TIMING_MM     4u   /* symbol timing by the reference graph's own serial tail -- clock_recovery_mm_ff(10,
                                      0.02296875, 0, 0.05, 0.005) -> binary_slicer_fb -> amps.recc with its buffer quirks, fed in
                                      256-byte work() calls (grc/ampsbs.grc:1751-1813, 1712-1750; lib/recc_impl.cc:93-145) --
                                      instead of the feed-forward detector.  One GPU thread walks the recurrence, so this mode
                                      runs at tens of Msymbols/s, not at the memory roofline.  In the burst record demod_index /
                                      sample_index are then nominal (recovered half-symbol index x 10), corr and run_length 0. */

#define AMPS_RX_INPUT_SC16    8u   /* samples arrive as interleaved int16 I,Q ("sc16": what the USRP puts on the wire before UHD
                                      converts to the fc32 stream of uhd.usrp_source, grc/ampsbs.grc:3750) through the *_sc16
                                      entry points: 4 bytes per sample over PCIe and out of HBM instead of 8; the conversion
                                      (float)int16 * sc16_scale happens in the front kernel.  Everything downstream is
                                      bit-identical to feeding amps_recc_iq_work() the converted floats. */

#define AMPS_RX_FUSED_SEARCH  16u  /* 10 MS/s: run the trigger search and the burst selection INSIDE the front kernel (two launches per
                                      call: front, capture) instead of as a launch of their own on a side stream (three; two for calls under 1.7 M samples).  Same
                                      results bit for bit.  Saves a launch; costs the front kernel its tail (the search of the last
                                      CTAs to finish cannot overlap anything), so the default keeps it outside -- DESIGN.md 4.2 has
                                      the measurements. */

typedef void (*amps_burst_cb)(const amps_burst *burst, void *user);

AMPS_B200_API int amps_recc_iq_create(const amps_recc_iq_params *params, amps_recc_iq **out);
AMPS_B200_API int amps_recc_iq_destroy(amps_recc_iq *h);
AMPS_B200_API int amps_recc_iq_reset(amps_recc_iq *h);          /* back to stream start (zero history) */

/* Host-buffer streaming call -- what a gr::sync_block::work() does with the scheduler's input
 * buffer (interleaved float re,im; nsamples complex samples).  Copies H2D, runs the fused kernels,
 * and invokes cb once per burst, in stream order (the equivalent of message_port_pub("bursts", ...),
 * lib/recc_impl.cc:126); the records arrive in a pinned host ring the capture kernel writes into.  Any
 * nsamples >= 0 is accepted; samples that do not fill a processing quantum (amps_recc_iq_granularity():
 * 1600 samples = 32 demodulated samples at 10 MS/s, 1536 at 400 kS/s) are carried to the next call.
 * A 10 MS/s call of 2^25 samples or more is uploaded in pieces of 2^24 samples on a stream of its own while the
 * kernels of the piece before run (same results; what is left to wait for after the last byte is one piece's kernels).
 * Returns AMPS_E_OVERFLOW ONCE (after delivering what was captured) when the device-side list of
 * undecided trigger candidates overflowed (> 8192) and candidates had to be dropped; the stream goes on. */
AMPS_B200_API int amps_recc_iq_work(amps_recc_iq *h, const float *iq_host, size_t nsamples,
                                    amps_burst_cb cb, void *user);

/* The same call for a handle created with AMPS_RX_INPUT_SC16: nsamples complex samples = 2*nsamples int16. */
AMPS_B200_API int amps_recc_iq_work_sc16(amps_recc_iq *h, const int16_t *iq_host, size_t nsamples,
                                         amps_burst_cb cb, void *user);

/* Device-resident variant: d_iq is a device pointer (16-byte aligned) on the handle's device; kernels are
 * enqueued and the call returns without synchronising: the front launch (filter + demod) on cuda_stream (a
 * cudaStream_t, NULL = default stream); the trigger search + burst selection and the capture on two streams of the
 * handle's own, ordered behind it by events, so that they overlap the front kernels of the next calls (up to four
 * calls deep; the call after that waits on the device, not on the host).  d_iq may be reused as soon as cuda_stream
 * has passed the call (its last samples are copied into the handle's history by the front kernel itself).
 * At 10 MS/s nsamples may be any count whose bytes are a multiple of 16 (fc32: even, sc16: multiple of 4):
 * what does not fill a 1600-sample quantum is carried, on the device, into the next call.  At 400 kS/s
 * nsamples must be a multiple of amps_recc_iq_granularity(). */
AMPS_B200_API int amps_recc_iq_submit_dev(amps_recc_iq *h, const void *d_iq, size_t nsamples, void *cuda_stream);
AMPS_B200_API int amps_recc_iq_submit_sc16_dev(amps_recc_iq *h, const void *d_iq, size_t nsamples, void *cuda_stream);
/* Waits for the stream, copies out the bursts published since the last collect (at most max; the
 * rest stays queued). */
AMPS_B200_API int amps_recc_iq_collect(amps_recc_iq *h, amps_burst *out, int max, int *n_out);
/* Zero-copy form: waits for the stream and exposes the pinned host ring the kernels publish into.
 * Bursts number first .. first+count-1 are at ring[(first + i) % ring_len]; call consume() when done. */
AMPS_B200_API int amps_recc_iq_peek(amps_recc_iq *h, const amps_burst **ring, uint32_t *ring_len,
                                    uint64_t *first, uint64_t *count);
AMPS_B200_API int amps_recc_iq_consume(amps_recc_iq *h, uint64_t count);
/* Non-blocking form of peek(): no stream synchronisation, reports the bursts the kernels have published so far
 * (a record is complete in host memory before the counter that covers it is written). */
AMPS_B200_API int amps_recc_iq_poll(amps_recc_iq *h, const amps_burst **ring, uint32_t *ring_len,
                                    uint64_t *first, uint64_t *count);
AMPS_B200_API int amps_recc_iq_granularity(const amps_recc_iq *h);
/* Inspection (tests / roofline accounting) */
AMPS_B200_API int amps_recc_iq_read_demod(amps_recc_iq *h, uint64_t first, float *out, size_t n);          /* 200 kS/s FM demod */
AMPS_B200_API int amps_recc_iq_read_baseband(amps_recc_iq *h, uint64_t first, float *out_iq, size_t n);    /* needs AMPS_RX_DUMP_BASEBAND */
AMPS_B200_API int amps_recc_iq_stats(const amps_recc_iq *h, uint64_t *samples_in, uint64_t *demod_out,
                                     uint64_t *bursts, uint64_t *kernel_launches);
/* Needs AMPS_RX_TIME_KERNELS: device time (ms) of the most recent front-end kernel launches, oldest
 * first, at most cap (the handle keeps the last 256).  Synchronises the stream. */
AMPS_B200_API int amps_recc_iq_front_times(amps_recc_iq *h, float *ms_out, int cap, int *n_out);
AMPS_B200_API int amps_recc_iq_get_taps(const amps_recc_iq *h, float *lpf_out, int cap);                   /* returns ntaps */
/* Measurement aid (handles created while AMPS_RX_PROF=1 is in the environment): 16 %globaltimer stamps (ns) per CTA of the
 * most recent 10 MS/s front launch -- tools/front_phases.py turns them into a phase breakdown. */
AMPS_B200_API int amps_recc_iq_debug_prof(amps_recc_iq *h, unsigned long long *out, int ctas);
/* Test aid, host arithmetic only: the front kernel's dealing of `tiles` 4800-sample tiles (nchan channels; equal_tiles > 0: they
 * all have that many) to at most `resident` CTAs.  *grid_out CTAs; lo_out[s] = first tile of CTA s for s = 0 .. grid (grid + 1
 * entries, at most 1025); owner_out[t] = the CTA that owns tile t (tiles entries). */
AMPS_B200_API int amps_b200_debug_deal(uint32_t tiles, uint32_t resident, uint32_t nchan, uint32_t equal_tiles, uint32_t *grid_out,
                                       uint32_t *lo_out, uint32_t *owner_out);

/* ------------------------------------------------------------------------------------------
 * Batched calls: K channels (handles) of ONE GPU served by one front launch + one capture launch per call
 * (per 64 channels).  This is "one carrier per GPU, round-robin beyond 8"  when there are more
 * carriers than GPUs, and the "one uploaded wideband buffer feeds several carriers" case: the reference
 * would instantiate K freq_xlating_fir_filter -> ... -> amps.recc chains on one uhd.usrp_source
 * (grc/ampsbs.grc:1814-1872 with K values of rx_offset, :212-238).
 * Handles must be fresh (or reset), 10 MS/s, feed-forward timing, same device and input format; while they
 * belong to a batch they are driven through it only.  Bursts are collected per handle with
 * amps_recc_iq_collect / _peek / _poll / _consume as usual.
 * ------------------------------------------------------------------------------------------ */
typedef struct amps_recc_iq_batch amps_recc_iq_batch;
typedef void (*amps_batch_burst_cb)(int channel, const amps_burst *burst, void *user);
AMPS_B200_API int amps_recc_iq_batch_create(amps_recc_iq *const *handles, int count, uint32_t flags /* AMPS_RX_TIME_KERNELS or 0 */,
                                            amps_recc_iq_batch **out);
AMPS_B200_API int amps_recc_iq_batch_destroy(amps_recc_iq_batch *b);     /* the handles stay valid and become individually usable again */
AMPS_B200_API int amps_recc_iq_batch_size(const amps_recc_iq_batch *b);
/* channel i gets nsamples[i] new samples at device pointer d_iq[i] (16-byte aligned, byte count a multiple of 16;
 * several channels may name the same buffer).  Returns without synchronising. */
AMPS_B200_API int amps_recc_iq_batch_submit_dev(amps_recc_iq_batch *b, const void *const *d_iq, const size_t *nsamples,
                                                void *cuda_stream);
/* ONE host buffer (fc32, or sc16 for sc16 handles), uploaded once; every channel demodulates it at its own
 * center_freq.  cb is invoked per burst with the channel index, channel by channel, in stream order within a channel. */
AMPS_B200_API int amps_recc_iq_batch_work_shared(amps_recc_iq_batch *b, const void *iq_host, size_t nsamples,
                                                 amps_batch_burst_cb cb, void *user);
/* needs AMPS_RX_TIME_KERNELS at batch_create: device time (ms) of the front launch(es) of the most recent calls */
AMPS_B200_API int amps_recc_iq_batch_front_times(amps_recc_iq_batch *b, float *ms_out, int cap, int *n_out);
AMPS_B200_API int amps_recc_iq_batch_stats(const amps_recc_iq_batch *b, uint64_t *calls, uint64_t *kernel_launches);

/* ------------------------------------------------------------------------------------------
 * recc_decode: message-only block (lib/recc_decode_impl.cc:81-169).  blob = 3374 hard half-symbols.
 * ------------------------------------------------------------------------------------------ */
typedef struct amps_recc_decode amps_recc_decode;
AMPS_B200_API int amps_recc_decode_create(int device, amps_recc_decode **out);
AMPS_B200_API int amps_recc_decode_destroy(amps_recc_decode *h);
AMPS_B200_API int amps_recc_decode_burst(amps_recc_decode *h, const uint8_t *blob3374, amps_recc_words *out);
/* batch form: nbursts blobs back to back */
AMPS_B200_API int amps_recc_decode_bursts(amps_recc_decode *h, const uint8_t *blobs, int nbursts, amps_recc_words *out);

/* ------------------------------------------------------------------------------------------
 * recc: byte-stream sink, compat mode (lib/recc_impl.cc:93-145 incl. its buffer quirks).
 * in = n hard half-symbols (0/1).  cb is invoked with each 3374-byte blob.
 * ------------------------------------------------------------------------------------------ */
typedef struct amps_recc amps_recc;
typedef void (*amps_blob_cb)(const uint8_t *blob3374, void *user);
AMPS_B200_API int amps_recc_create(int device, amps_recc **out);
AMPS_B200_API int amps_recc_destroy(amps_recc *h);
/* one work() call; returns AMPS_OK (the reference returns 0 items and consumes n) */
AMPS_B200_API int amps_recc_work(amps_recc *h, const uint8_t *in, int n, amps_blob_cb cb, void *user);
/* a whole schedule of work() calls in one launch: chunk_sizes[nchunks], in = concatenated chunks */
AMPS_B200_API int amps_recc_work_chunks(amps_recc *h, const uint8_t *in, const int *chunk_sizes, int nchunks,
                                        amps_blob_cb cb, void *user);

/* ------------------------------------------------------------------------------------------
 * focc: FOCC Manchester half-symbol source (lib/focc_impl.cc:104-136, 486-647).
 * ------------------------------------------------------------------------------------------ */
typedef struct amps_focc amps_focc;
AMPS_B200_API int amps_focc_create(unsigned long symrate, int aggressive_registration, int device, amps_focc **out);
AMPS_B200_API int amps_focc_destroy(amps_focc *h);
/* work(): writes up to noutput_items bytes (+1 = 0x01, -1 = 0xFF) to out (host memory) and returns
 * the number produced in *produced: at most one 23/22-bit burst per call, possibly 0; -1 (WORK_DONE)
 * when noutput_items < 1 (lib/focc_impl.cc:590-593,630-632). */
AMPS_B200_API int amps_focc_work(amps_focc *h, uint8_t *out, int noutput_items, int *produced);
/* bulk form: the concatenation of successive work() outputs until exactly n bytes were produced */
AMPS_B200_API int amps_focc_generate(amps_focc *h, uint8_t *out, size_t n);
AMPS_B200_API int amps_focc_generate_dev(amps_focc *h, void *d_out, size_t n, void *cuda_stream);
/* the same stream as DATA BITS (one byte per bit, busy/idle resolved): the input of amps_fwd_*_bits; only valid on a
 * bit boundary; advances the state exactly as generating nbits * 2 * (symrate / 20000) bytes would */
AMPS_B200_API int amps_focc_generate_bits(amps_focc *h, uint8_t *out, size_t nbits);
AMPS_B200_API int amps_focc_generate_bits_dev(amps_focc *h, void *d_out, size_t nbits, void *cuda_stream);
/* focc_words message (lib/focc_impl.cc:521-563): stream 1=A 2=B 3=BOTH, words28 = nwords x 28 bytes */
AMPS_B200_API int amps_focc_push_words(amps_focc *h, long stream, const uint8_t *words28, long nwords);
AMPS_B200_API int amps_focc_set_busy_idle(amps_focc *h, int idle);      /* lib/amps_common.h:7 */

/* ------------------------------------------------------------------------------------------
 * fvc: FVC blank-and-burst source (lib/fvc_impl.cc:56-193).
 * ------------------------------------------------------------------------------------------ */
typedef struct amps_fvc amps_fvc;
AMPS_B200_API int amps_fvc_create(unsigned long symrate, int device, amps_fvc **out);
AMPS_B200_API int amps_fvc_destroy(amps_fvc *h);
/* fvc_words message (lib/fvc_impl.cc:109-143); has_timer/timer = the optional trailing uint64 */
AMPS_B200_API int amps_fvc_push_words(amps_fvc *h, const uint8_t *words28, long nwords, int has_timer, uint64_t timer);
/* work(): *produced = items produced; while no word was ever pushed it is noutput_items and `out` is
 * NOT written, as in the reference (lib/fvc_impl.cc:159-161).  *fvc_off is set when the "fvc off" PDU is
 * due (:163-171). */
AMPS_B200_API int amps_fvc_work(amps_fvc *h, uint8_t *out, int noutput_items, int *produced, int *fvc_off);
/* the same replay as DATA BITS (one byte per bit; 0xFF = muted while no word was ever pushed), for amps_fwd_*_bits */
AMPS_B200_API int amps_fvc_work_bits(amps_fvc *h, uint8_t *out_bits, int nbits, int *produced, int *fvc_off);

/* ------------------------------------------------------------------------------------------
 * Fused forward path: symbols -> char_to_float -> frequency_modulator_fc -> pfb interpolator (x4, the
 * reference's taps @400 kS/s) -> CIC^3 x25 -> mix -> sum -> x0.5 (grc/ampsbs.grc:1159-1252, 574-659,
 * 2120-2229, 817-942, 1006-1056, 1355-1405) at 10 MS/s output.  Up to 3 carriers (FOCC @0 Hz + two
 * FVC legs).  A symbol byte of 0 mutes that symbol (mute_xx, :1508-1601).
 * ------------------------------------------------------------------------------------------ */
typedef struct amps_fwd amps_fwd;
typedef struct amps_fwd_params {
    double   samp_rate;          /* output rate, 10e6 */
    double   symrate;            /* symbol-stream rate feeding the FM modulator, 100e3 (grc/ampsbs.grc:135,317) */
    double   max_deviation;      /* 8000 (grc/ampsbs.grc:209) */
    int      device;
    int      ncarriers;          /* 1..3 */
    double   carrier_freq[3];    /* 0, 60e3, 90e3 (grc/ampsbs.grc:841,904) */
    double   lpf_transition[3];  /* x4 pfb interpolator taps firdes.low_pass(1, 400e3, 10e3, tw): 5e3 FOCC, 3e3 FVC (:2227,:2172) */
    float    out_scale;          /* 0.5 (:1367) */
    uint32_t max_samples;
} amps_fwd_params;
AMPS_B200_API int amps_fwd_create(const amps_fwd_params *p, amps_fwd **out);
AMPS_B200_API int amps_fwd_destroy(amps_fwd *h);
AMPS_B200_API int amps_fwd_reset(amps_fwd *h);
/* sym[c] = host arrays of nsym +1/-1 (0x01/0xFF, 0 = muted) bytes per carrier; out_iq_host gets
 * nsym * (samp_rate/symrate) complex samples. */
AMPS_B200_API int amps_fwd_work(amps_fwd *h, const uint8_t *const *sym, size_t nsym, float *out_iq_host);
AMPS_B200_API int amps_fwd_submit_dev(amps_fwd *h, const void *const *d_sym, size_t nsym, void *d_out_iq, void *cuda_stream);
/* Manchester-bit fast path: the same chain driven by DATA BITS, one byte per 10 kbit/s bit (0, 1, 0xFF = muted;
 * bit 0 -> half-symbols (+1,-1), bit 1 -> (-1,+1) as lib/amps_packet.h:52-70), 1000 output samples per bit.  Because the
 * FM phase returns to zero at every Manchester bit boundary the interpolator output is a sum of per-bit table
 * entries: same result as feeding the expanded half-symbols to amps_fwd_work (within fp32 rounding), several times
 * faster.  A handle streams either half-symbols or bits; reset() to switch. */
AMPS_B200_API int amps_fwd_work_bits(amps_fwd *h, const uint8_t *const *bits, size_t nbits, float *out_iq_host);
AMPS_B200_API int amps_fwd_submit_bits_dev(amps_fwd *h, const void *const *d_bits, size_t nbits, void *d_out_iq, void *cuda_stream);
/* Voice legs: wavfile/audio @16 kS/s + 6 kHz SAT -> analog.nbfm_tx(16000, 16000, tau 75 us, max_dev 8 kHz)
 * -> [mute_xx audio_mute] -> pfb.arb_resampler_ccf(25, voice_lpf_taps = firdes.low_pass(3, 400e3, 15e3, 6e3, BLACKMAN), 8 arms)
 * added to a carrier's 400 kS/s samples in front of its mixer (grc/ampsbs.grc:715-773, 943-1005, 1994-2119, 4494-4500,
 * 4632-4638).  In the reference graph the gated leg shares the +60 kHz mixer with the FVC data and the open leg feeds the
 * +90 kHz mixer alone (give that carrier an all-zero symbol stream). */
typedef struct amps_fwd_voice_params {
    int    carrier_gated;        /* carrier index of the leg behind mute_xx(audio_mute) (1 in the reference graph), -1 = none */
    int    carrier_open;         /* carrier index of the leg that is always on (2 in the reference graph), -1 = none */
    double audio_rate;           /* 16000 (x25 to the reference's 400 kS/s) */
    double max_dev;              /* 8e3 */
    double tau;                  /* 75e-6 */
    double sat_freq;             /* 6000 */
    double sat_amp;              /* 0.05; 0 = no supervisory tone */
} amps_fwd_voice_params;
AMPS_B200_API int amps_fwd_enable_voice(amps_fwd *h, const amps_fwd_voice_params *vp);
/* half-symbol streams as amps_fwd_work plus 4 nsym / 25 audio samples (float, 16 kS/s); nsym must be a multiple of 25.
 * audio_mute != 0 mutes the gated leg for this call's audio samples. */
AMPS_B200_API int amps_fwd_work_voice(amps_fwd *h, const uint8_t *const *sym, const float *audio, size_t nsym, int audio_mute,
                                      float *out_iq_host);
AMPS_B200_API int amps_fwd_submit_voice_dev(amps_fwd *h, const void *const *d_sym, const void *d_audio, size_t nsym,
                                            int audio_mute, void *d_out_iq, void *cuda_stream);
AMPS_B200_API int amps_fwd_interp(const amps_fwd *h);
AMPS_B200_API int amps_fwd_get_taps(const amps_fwd *h, int carrier, float *out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* AMPS_B200_H */
