/*
 * focc_src.c -- ORACLE (test infrastructure).  FOCC Manchester half-symbol source.
 * Restates lib/focc_impl.cc:104-136 (ctor), :178-218 (frame layout), :383-480 (superframes),
 * :486-519 (segment/frame advance), :521-580 (word injection), :582-647 (work), and the
 * Manchester/oversample encoding of lib/amps_packet.h:47-76.
 *
 * Representation differs from the reference on purpose (no pre-encoded segments): a frame is
 * 463 bit slots; the byte at (bit b, offset o) is computed on the fly.  Behaviour kept:
 *   - bit 0 -> (+1 x sps, -1 x sps), bit 1 -> (-1 x sps, +1 x sps), -1 stored as 0xFF;
 *   - busy/idle slots take the CURRENT value of the B/I flag at emission time (:606-610);
 *   - an END marker follows bit 22 and then every 22 bits; work() returns when it steps over one
 *     (:630-632), so a call yields at most one 23/22-bit burst and may return 0;
 *   - the frame index advances (and a queued frame may replace a filler slot) while stepping
 *     over the END marker that closes a frame (:491-507).
 */
#include "amps_oracle.h"
#include <stdlib.h>
#include <string.h>

#define FRAME_BITS 463
#define SLOT_BI 2

typedef struct frame {
    uint8_t slot[FRAME_BITS];   /* 0, 1 or SLOT_BI */
    int     ephemeral, filler;
    struct frame *next;         /* queue link */
} frame;

struct orc_focc {
    unsigned sps;
    int      nframes;
    frame   *super[40];
    frame   *cur;
    int      frame_idx;
    int      bit;               /* next bit slot to emit, 0..462 (463 == at closing END) */
    int      off;               /* samples of that bit already emitted, 0..2*sps-1 */
    int      at_end;            /* sitting on an END marker */
    int      bi;                /* busy_idle_bit (lib/amps_common.h:7), 1 = idle */
    frame   *qhead, *qtail;
};

static frame *make_frame(const uint8_t wa[28], const uint8_t wb[28], int ephemeral, int filler) {
    static const uint8_t dot[10] = {1,0,1,0,1,0,1,0,1,0};
    static const uint8_t sync[11] = {1,1,1,0,0,0,1,0,0,1,0};
    frame *f = (frame *)calloc(1, sizeof *f);
    uint8_t ea[40], eb[40];
    orc_bch_encode_40_28(wa, ea);
    orc_bch_encode_40_28(wb, eb);
    int p = 0;
    f->slot[p++] = SLOT_BI; memcpy(f->slot + p, dot, 10); p += 10;
    f->slot[p++] = SLOT_BI; memcpy(f->slot + p, sync, 11); p += 11;
    for (int r = 0; r < 5; r++) {
        for (int w = 0; w < 2; w++) {
            const uint8_t *e = w ? eb : ea;
            for (int q = 0; q < 4; q++) {
                f->slot[p++] = SLOT_BI;
                memcpy(f->slot + p, e + 10 * q, 10); p += 10;
            }
        }
    }
    /* p == 463 (lib/focc_impl.cc:246) */
    f->ephemeral = ephemeral; f->filler = filler;
    return f;
}

static int is_end_after(int bit_index) { /* END marker sits after this bit? */
    int n = bit_index + 1;
    return n >= 23 && (n - 23) % 22 == 0;
}

static void add2(orc_focc *f, const uint8_t w[28], int filler) {
    f->super[f->nframes++] = make_frame(w, w, 0, filler);
}

orc_focc *orc_focc_new(unsigned long symrate, int aggressive) {
    orc_focc *f = (orc_focc *)calloc(1, sizeof *f);
    uint8_t w[28];
    f->sps = (unsigned)(symrate / 20000);
    f->bi = 1;
    int halves = aggressive ? 2 : 1;
    for (int h = 0; h < halves; h++) {
        orc_overhead_word_1(w, 0, 16, 1, 0, 0, aggressive ? 4 : 3); add2(f, w, 0);
        orc_overhead_word_2(w, 0, 1, 1, 1, 1, 0, 23, 1, 1, 23, 0);  add2(f, w, 0);
        orc_access_type_global_action(w, 0, 0);                     add2(f, w, 0);
        if (aggressive) { orc_reg_increment_global_action(w, 0, 100, 0); add2(f, w, 0); }
        orc_registration_id(w, 0, (aggressive && h == 1) ? 500 : 0, 1); add2(f, w, 0);
        orc_control_filler_word(w);
        int nfill = aggressive ? 14 : 15;
        for (int i = 0; i < nfill; i++) add2(f, w, 1);
    }
    f->frame_idx = 0;
    f->cur = f->super[0];
    f->bit = 0; f->off = 0; f->at_end = 0;
    return f;
}

void orc_focc_free(orc_focc *f) {
    if (!f) return;
    if (f->cur && f->cur->ephemeral) free(f->cur);
    for (int i = 0; i < f->nframes; i++) free(f->super[i]);
    while (f->qhead) { frame *n = f->qhead->next; free(f->qhead); f->qhead = n; }
    free(f);
}

int orc_focc_superframe_frames(const orc_focc *f) { return f->nframes; }

int orc_focc_push_words(orc_focc *f, long stream, const uint8_t *words28, long nwords) {
    uint8_t fill[28];
    orc_control_filler_word(fill);
    for (long i = 0; i < nwords; i++) {
        const uint8_t *w = words28 + 28 * i;
        frame *fr;
        if (stream == 1) fr = make_frame(w, fill, 1, 0);
        else if (stream == 2) fr = make_frame(fill, w, 1, 0);
        else if (stream == 3) fr = make_frame(w, w, 1, 0);
        else return -1;
        fr->next = NULL;
        if (f->qtail) f->qtail->next = fr; else f->qhead = fr;
        f->qtail = fr;
    }
    return 0;
}

static void step_over_end(orc_focc *f) {
    if (f->bit == FRAME_BITS) { /* END that closes the frame: advance frame, maybe pop the queue */
        f->frame_idx = (f->frame_idx + 1) % f->nframes;
        if (f->cur->ephemeral) free(f->cur);
        f->cur = f->super[f->frame_idx];
        if (f->cur->filler && f->qhead) {
            frame *q = f->qhead;
            f->qhead = q->next;
            if (!f->qhead) f->qtail = NULL;
            f->cur = q;
        }
        f->bit = 0;
    }
    f->at_end = 0;
}

int orc_focc_work(orc_focc *f, uint8_t *out, int n) {
    if (n < 1) return -1;
    const int two = 2 * (int)f->sps;
    int produced = 0;
    while (produced < n) {
        if (f->at_end) { step_over_end(f); return produced; }
        unsigned s = f->cur->slot[f->bit];
        unsigned bitval = (s == SLOT_BI) ? (unsigned)(f->bi ? 1 : 0) : s;
        int take = two - f->off;
        if (take > n - produced) take = n - produced;
        for (int i = 0; i < take; i++) {
            int o = f->off + i;
            int first_half = o < (int)f->sps;
            /* bit 1: (-1,+1); bit 0: (+1,-1) */
            int high = bitval ? !first_half : first_half;
            out[produced + i] = high ? 0x01 : 0xFF;
        }
        produced += take;
        f->off += take;
        if (f->off == two) {
            f->off = 0;
            if (is_end_after(f->bit)) f->at_end = 1;
            f->bit++;
        }
    }
    return produced;
}
