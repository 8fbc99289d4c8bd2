/*
 * recc_capture.c -- ORACLE (test infrastructure).  RECC trigger search + 3374-symbol capture on
 * hard half-symbols, compat mode: restates lib/recc_impl.cc:51-65 (Manchester polarity),
 * :67-83 (ctor, trigger string), :93-145 (work) INCLUDING the buffer quirks:
 *   - wrap copies bytes [61440,65536) of the buffer (capacity tail, not data tail) and forgets a
 *     pending trigger (:104-108);
 *   - only the last min(len, n+73) bytes are searched, first match wins (:115-119);
 *   - publish when strictly more than 3374 symbols follow the trigger (:124-126);
 *   - after a publish the LAST `startoff` bytes are moved to the front and len shrinks by
 *     startoff (:129-134).
 */
#include "amps_oracle.h"
#include <stdlib.h>
#include <string.h>

#define BUFSZ 65536
#define WINDOW 4096

struct orc_recc {
    uint8_t buf[BUFSZ];
    size_t  len;
    long    pending;      /* offset of a found trigger, -1 if none */
    uint8_t trig[ORC_RECC_TRIGGER_LEN];
};

void orc_recc_trigger(uint8_t out74[74]) {
    /* 26 dotting bits "10"x13 then word sync 11100010010; bit 0 -> (1,0), bit 1 -> (0,1) */
    static const char *sync = "11100010010";
    int o = 0;
    for (int i = 0; i < 37; i++) {
        int bit = (i < 26) ? ((i & 1) ^ 1) : (sync[i - 26] - '0');
        out74[o++] = bit ? 0 : 1;
        out74[o++] = bit ? 1 : 0;
    }
}

orc_recc *orc_recc_new(void) {
    orc_recc *r = (orc_recc *)calloc(1, sizeof *r);
    r->pending = -1;
    orc_recc_trigger(r->trig);
    return r;
}
void orc_recc_free(orc_recc *r) { free(r); }
size_t orc_recc_buflen(const orc_recc *r) { return r->len; }

int orc_recc_work(orc_recc *r, const uint8_t *in, int n, orc_burst_cb cb, void *user) {
    if (n < 1) return 0;
    if (n >= BUFSZ - WINDOW) return -2;          /* the reference asserts (:103) */
    if (r->len + (size_t)n > BUFSZ) {
        memmove(r->buf, r->buf + (BUFSZ - WINDOW), WINDOW);
        r->len = WINDOW;
        r->pending = -1;
    }
    memmove(r->buf + r->len, in, (size_t)n);
    r->len += (size_t)n;
    if (r->len > ORC_RECC_TRIGGER_LEN) {
        size_t searchsz = (size_t)n + ORC_RECC_TRIGGER_LEN - 1;
        if (searchsz > r->len) searchsz = r->len;
        if (r->pending < 0) {
            size_t base = r->len - searchsz;
            for (size_t i = 0; i + ORC_RECC_TRIGGER_LEN <= searchsz; i++) {
                if (memcmp(r->buf + base + i, r->trig, ORC_RECC_TRIGGER_LEN) == 0) {
                    r->pending = (long)(base + i);
                    break;
                }
            }
        }
        if (r->pending >= 0) {
            size_t start = (size_t)r->pending;
            long captured = (long)r->len - (long)start - ORC_RECC_TRIGGER_LEN;
            if (captured > ORC_RECC_CAPTURE_LEN) {
                if (cb) cb(r->buf + start + ORC_RECC_TRIGGER_LEN, user);
                size_t tomove = start;           /* == len - (captured + trigger_len) */
                if (tomove > 0) memmove(r->buf, r->buf + (r->len - tomove), tomove);
                r->len -= tomove;
                r->pending = -1;
            }
        }
    }
    return 0;
}
