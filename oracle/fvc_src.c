/*
 * fvc_src.c -- ORACLE (test infrastructure).  FVC blank-and-burst half-symbol source.
 * Restates lib/fvc_impl.cc:56-67 (ctor), :71-88 (Manchester + oversample), :98-107 (BCH),
 * :109-143 (word train), :152-193 (work / replay / timerhack).
 *
 * The reference replays a std::queue<bool> copy of everything ever pushed; here the train is a
 * flat bit array and a read cursor.  Behaviour kept: per word 101-bit dotting, then 11 x
 * (11-bit word sync + 40-bit BCH word) with 37-bit dotting between repeats (1032 bits); the
 * train only ever grows; each work() call emits at most the remainder of the current replay;
 * with no word ever pushed work() returns n and writes nothing (:159-161); timerhack counts
 * down once per replay start and raises "fvc off" when it reaches zero (:163-171).
 */
#include "amps_oracle.h"
#include <stdlib.h>
#include <string.h>

struct orc_fvc {
    unsigned sps;
    uint8_t *bits;       /* accumulated train, one byte per bit */
    size_t   nbits, cap;
    size_t   replay_len; /* length in bytes of the replay snapshot being emitted (0 = none) */
    size_t   replay_pos; /* bytes already emitted from the snapshot */
    uint64_t timer;
};

orc_fvc *orc_fvc_new(unsigned long symrate) {
    orc_fvc *f = (orc_fvc *)calloc(1, sizeof *f);
    f->sps = (unsigned)(symrate / 20000);
    return f;
}
void orc_fvc_free(orc_fvc *f) { if (f) { free(f->bits); free(f); } }

static void put(orc_fvc *f, const uint8_t *b, size_t n) {
    if (f->nbits + n > f->cap) {
        f->cap = (f->nbits + n) * 2 + 1024;
        f->bits = (uint8_t *)realloc(f->bits, f->cap);
    }
    memcpy(f->bits + f->nbits, b, n);
    f->nbits += n;
}

int orc_fvc_push_words(orc_fvc *f, const uint8_t *words28, long nwords, int has_timer, uint64_t timer) {
    static const uint8_t sync[11] = {1,1,1,0,0,0,1,0,0,1,0};
    uint8_t dot[101];
    for (int i = 0; i < 101; i++) dot[i] = (uint8_t)((i & 1) ^ 1);   /* 1010...1 */
    if (has_timer) f->timer = timer;
    for (long w = 0; w < nwords; w++) {
        uint8_t enc[40];
        orc_bch_encode_40_28(words28 + 28 * w, enc);
        put(f, dot, 101);
        for (int j = 0; j < 11; j++) {
            put(f, sync, 11);
            put(f, enc, 40);
            if (j < 10) put(f, dot, 37);
        }
    }
    return 0;
}

int orc_fvc_work(orc_fvc *f, uint8_t *out, int n, int *fvc_off) {
    if (fvc_off) *fvc_off = 0;
    if (f->nbits == 0) return n;                 /* claims n items, writes none */
    if (f->replay_pos == f->replay_len) {        /* replay queue empty: take a fresh snapshot */
        if (f->timer >= 1) {
            f->timer--;
            if (f->timer == 0 && fvc_off) *fvc_off = 1;
        }
        f->replay_len = f->nbits * 2 * f->sps;
        f->replay_pos = 0;
    }
    size_t left = f->replay_len - f->replay_pos;
    int take = (size_t)n < left ? n : (int)left;
    for (int i = 0; i < take; i++) {
        size_t p = f->replay_pos + (size_t)i;
        size_t bit = p / (2 * f->sps);
        int first_half = (p % (2 * f->sps)) < f->sps;
        int high = f->bits[bit] ? !first_half : first_half;   /* bit 1 -> (low, high) */
        out[i] = high ? 0x01 : 0xFF;
    }
    f->replay_pos += (size_t)take;
    return take;
}
