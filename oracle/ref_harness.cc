// ref_harness.cc -- TEST INFRASTRUCTURE.  C entry points around the UNMODIFIED gr-amps block sources, which
// oracle/Makefile (target _ref) compiles from where they lie under /root/reference/lib against stand-in headers for
// their absent third-party dependencies (oracle/ref_shim: GNU Radio block base classes + PMT, Boost, IT++).  What
// runs behind these functions is the reference's own focc_impl / fvc_impl / recc_impl / recc_decode_impl /
// command_processor_impl / amps_packet / utils code; only the scheduler, the message transport and itpp::BCH are ours.
// Used by tests/ (to pin oracle/*.c and to generate tests/golden/ref_*.json) -- never by the product.
//
// Result structs are the oracle's own (amps_oracle.h: layouts only, nothing from liboracle.so is linked), so the two
// can be compared field by field.
#include <fcntl.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include <queue>
#include <sstream>
#include <fstream>
#include <algorithm>
#include <itpp/comm/bch.h>
#include <gnuradio/sync_block.h>
#include <amps/recc_decode.h>
#include "amps_packet.h"
#define private public   /* recc_decode_impl::recc_bch_decode is private; the harness needs the per-word validity */
#include "recc_decode_impl.h"
#undef private
#include "command_processor_impl.h"
#include "focc_impl.h"
#include "fvc_impl.h"
#include "recc_impl.h"
#include "utils.h"

#include "amps_oracle.h"

namespace gr { namespace amps { void manchester_encode(const char *src, const size_t srclen, unsigned char *dst); } }   /* lib/recc_impl.cc:51 */
using namespace gr::amps;

namespace {

// The reference prints on every capture / decode (lib/recc_impl.cc:128, lib/recc_decode_impl.cc:63-70,85-89).  Send
// fd 1 to a scratch file for the duration of a call and hand the text back (the LOG_DEBUG lines carry MIN/ESN/dialled
// digits, which are not published on any message port).
class stdout_capture {
public:
    stdout_capture() {
        fflush(stdout);
        std::cout.flush();
        d_saved = dup(1);
        char path[] = "/tmp/amps_ref_XXXXXX";
        d_fd = mkstemp(path);
        unlink(path);
        dup2(d_fd, 1);
    }
    std::string finish() {
        if (d_saved < 0) return d_text;
        fflush(stdout);
        std::cout.flush();
        dup2(d_saved, 1);
        close(d_saved);
        d_saved = -1;
        off_t n = lseek(d_fd, 0, SEEK_END);
        d_text.resize(n > 0 ? (size_t)n : 0);
        if (n > 0 && pread(d_fd, &d_text[0], (size_t)n, 0) != n) d_text.clear();
        close(d_fd);
        return d_text;
    }
    ~stdout_capture() { finish(); }
private:
    int d_saved, d_fd;
    std::string d_text;
};

struct recorded { std::string port; pmt::pmt_t msg; };

// receives whatever a block publishes, in order
class recorder : public gr::block {
public:
    recorder() : gr::block("recorder", gr::io_signature::make(0, 0, 0), gr::io_signature::make(0, 0, 0)) {}
    void listen(gr::basic_block &src, const std::string &port) {
        message_port_register_in(pmt::mp(port));
        set_msg_handler(pmt::mp(port), [this, port](pmt::pmt_t m) { recorded r; r.port = port; r.msg = m; log.push_back(r); });
        gr::msg_connect(src, port, *this, port);
    }
    std::vector<recorded> log;
};

std::string pdu_text(const pmt::pmt_t &pdu) {
    size_t n = 0;
    const uint8_t *p = pmt::u8vector_elements(pmt::cdr(pdu), n);
    return std::string((const char *)p, n);
}
void put_text(char *dst, size_t cap, const std::string &s) {
    size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(dst, s.data(), n);
    dst[n] = 0;
}

struct focc_h { focc::sptr b; };
struct fvc_h { fvc::sptr b; recorder rec; };
struct recc_h { recc::sptr b; recorder rec; };

}  // namespace

extern "C" {
#define REF_API __attribute__((visibility("default")))

/* ---- amps.focc (lib/focc_impl.cc) */
REF_API void *ref_focc_new(unsigned long symrate, int aggressive) {
    stdout_capture q;
    focc_h *h = new focc_h();
    h->b = focc::make(symrate, aggressive != 0);
    return h;
}
REF_API void ref_focc_free(void *p) { stdout_capture q; delete (focc_h *)p; }
REF_API int ref_focc_work(void *p, uint8_t *out, int n) {
    stdout_capture q;
    gr_vector_const_void_star in;
    gr_vector_void_star outs(1);
    outs[0] = out;
    return ((focc_h *)p)->b->work(n, in, outs);
}
REF_API void ref_focc_push_words(void *p, long stream, const uint8_t *words28, long nwords) {
    stdout_capture q;
    std::vector<pmt::pmt_t> items;
    items.push_back(pmt::from_long(stream));
    items.push_back(pmt::from_long(nwords));
    for (long i = 0; i < nwords; ++i) items.push_back(pmt::mp((const void *)(words28 + 28 * i), 28));
    ((focc_h *)p)->b->dispatch_msg("focc_words", pmt::make_tuple_v(items));
}
REF_API void ref_set_busy_idle(int bit) { busy_idle_bit = bit != 0; }   /* lib/amps_common.h:7 */

/* ---- amps.fvc (lib/fvc_impl.cc) */
REF_API void *ref_fvc_new(unsigned long symrate) {
    stdout_capture q;
    fvc_h *h = new fvc_h();
    h->b = fvc::make(symrate);
    h->rec.listen(*h->b, "command_out");
    return h;
}
REF_API void ref_fvc_free(void *p) { stdout_capture q; delete (fvc_h *)p; }
REF_API void ref_fvc_push_words(void *p, const uint8_t *words28, long nwords, int has_timer, uint64_t timer) {
    stdout_capture q;
    std::vector<pmt::pmt_t> items;
    items.push_back(pmt::from_long(nwords));
    for (long i = 0; i < nwords; ++i) items.push_back(pmt::mp((const void *)(words28 + 28 * i), 28));
    if (has_timer) items.push_back(pmt::from_uint64(timer));
    ((fvc_h *)p)->b->dispatch_msg("fvc_words", pmt::make_tuple_v(items));
}
/* *fvc_off = number of "fvc off" PDUs published on command_out during this call */
REF_API int ref_fvc_work(void *p, uint8_t *out, int n, int *fvc_off) {
    stdout_capture q;
    fvc_h *h = (fvc_h *)p;
    h->rec.log.clear();
    gr_vector_const_void_star in(1);
    in[0] = NULL;
    gr_vector_void_star outs(1);
    outs[0] = out;
    int r = h->b->work(n, in, outs);
    int off = 0;
    for (size_t i = 0; i < h->rec.log.size(); ++i) off += pdu_text(h->rec.log[i].msg) == "fvc off";
    if (fvc_off) *fvc_off = off;
    return r;
}

/* ---- amps.recc (lib/recc_impl.cc) */
typedef void (*ref_burst_cb)(const uint8_t *blob3374, void *user);
REF_API void *ref_recc_new(void) {
    stdout_capture q;
    recc_h *h = new recc_h();
    h->b = recc::make();
    h->rec.listen(*h->b, "bursts");
    return h;
}
REF_API void ref_recc_free(void *p) { stdout_capture q; delete (recc_h *)p; }
REF_API int ref_recc_work(void *p, const uint8_t *in, int n, ref_burst_cb cb, void *user) {
    recc_h *h = (recc_h *)p;
    int r;
    {
        stdout_capture q;
        h->rec.log.clear();
        gr_vector_const_void_star ins(1);
        ins[0] = in;
        gr_vector_void_star outs;
        r = h->b->work(n, ins, outs);
    }
    for (size_t i = 0; i < h->rec.log.size(); ++i)
        if (cb && pmt::blob_length(h->rec.log[i].msg) == 3374) cb((const uint8_t *)pmt::blob_data(h->rec.log[i].msg), user);
    return r;
}
REF_API void ref_recc_trigger(uint8_t out74[74]) {      /* lib/recc_impl.cc:51-65,76-79 */
    const char *trigbuf = "1010101010101010101010101011100010010";
    manchester_encode(trigbuf, strlen(trigbuf), out74);
}

/* ---- amps.recc_decode (lib/recc_decode_impl.cc) */
static void fill_actions(const std::vector<recorded> &log, orc_recc_actions *a) {
    memset(a, 0, sizeof *a);
    a->fvc_mute = a->audio_mute = -1;
    for (size_t i = 0; i < log.size(); ++i) {
        const recorded &r = log[i];
        if (r.port == "focc_words") {
            a->focc_stream = pmt::to_long(pmt::tuple_ref(r.msg, 0));
            a->n_focc = (int32_t)pmt::to_long(pmt::tuple_ref(r.msg, 1));
            for (int w = 0; w < a->n_focc && w < 2; ++w) memcpy(a->focc_words[w], pmt::blob_data(pmt::tuple_ref(r.msg, 2 + w)), 28);
        } else if (r.port == "fvc_words") {
            a->has_fvc = 1;
            long nw = pmt::to_long(pmt::tuple_ref(r.msg, 0));
            memcpy(a->fvc_word, pmt::blob_data(pmt::tuple_ref(r.msg, 1)), 28);
            if ((long)pmt::length(r.msg) > 1 + nw) a->fvc_timer = pmt::to_uint64(pmt::tuple_ref(r.msg, 1 + nw));
        } else if (r.port == "fvc_mute") {
            a->fvc_mute = pmt::to_bool(r.msg);
        } else if (r.port == "audio_mute") {
            a->audio_mute = pmt::to_bool(r.msg);
        } else if (r.port == "command_out") {
            put_text(a->command, sizeof a->command, pdu_text(r.msg));
        }
    }
}
/* runs recc_decode_impl::bursts_message on one blob; everything it publishes -> *a, everything it prints -> log */
REF_API void ref_recc_bursts_message(const uint8_t *blob, size_t len, orc_recc_actions *a, char *log, size_t logcap) {
    stdout_capture q;
    recc_decode::sptr b = recc_decode::make();
    recorder rec;
    const char *ports[] = {"focc_words", "fvc_words", "audio_mute", "fvc_mute", "command_out"};
    for (int i = 0; i < 5; ++i) rec.listen(*b, ports[i]);
    b->dispatch_msg("bursts", pmt::mp((const void *)blob, len));
    fill_actions(rec.log, a);
    std::string text = q.finish();
    if (log && logcap) put_text(log, logcap, text);
}
/* Manchester decode, per-word BCH validity and the word parsers of lib/amps_packet.h, called the way bursts_message
 * calls them (lib/recc_decode_impl.cc:90-117, 130, 146, 161); kind/esn/min/dialed are left zero (they exist only in
 * the log of ref_recc_bursts_message). */
REF_API void ref_recc_fields(const uint8_t blob[3374], orc_recc_result *o) {
    stdout_capture q;
    memset(o, 0, sizeof *o);
    recc_decode_impl dec;
    o->dcc_errs = (uint8_t)manchester_decode_binbuf(blob, o->dcc, 7);
    unsigned char decword[48];
    for (int i = 0; i < 7; ++i) o->errs[i] = (uint16_t)manchester_decode_binbuf(&blob[14 + 480 * i], o->words[i], 240);
    for (int w = 0; w < 7; ++w) {
        o->valid_repeat[w] = 5;
        for (int r = 0; r < 5; ++r) {
            o->valid[w] = dec.recc_bch_decode(&o->words[w][r * 48], decword);
            if (o->valid[w]) { o->valid_repeat[w] = (uint8_t)r; break; }
        }
    }
    recc_word_a wa(o->words[0]);
    o->F = wa.F; o->NAWC = wa.NAWC; o->T = wa.T; o->S = wa.S; o->E = wa.E; o->ER = wa.ER; o->SCM = wa.SCM;
    o->MIN1 = (uint32_t)wa.MIN1;
    recc_word_b wb(o->words[1]);
    o->B_F = wb.F; o->B_NAWC = wb.NAWC; o->MSG_TYPE = wb.MSG_TYPE; o->ORDQ = wb.ORDQ; o->ORDER = wb.ORDER;
    o->LT = wb.LT; o->EP = wb.EP; o->SCM4 = wb.SCM4; o->MPCI = wb.MPCI; o->SDCC1 = wb.SDCC1; o->SDCC2 = wb.SDCC2;
    o->MIN2 = (uint16_t)wb.MIN2;
    recc_word_c_serial wc(o->words[2]);
    o->word_c_serial = (uint32_t)wc.SERIAL;
    put_text(o->min, sizeof o->min, calc_min(wa, wb));
}
REF_API void ref_called_digits(const uint8_t word48[48], char out[16]) {     /* lib/amps_packet.h:200-262 */
    stdout_capture q;
    recc_word_called w(word48);
    put_text(out, 16, w.digits());
}

/* ---- amps.command_processor (lib/command_processor_impl.cc) */
REF_API void ref_command(const char *cmd, orc_cmd_actions *a) {
    stdout_capture q;
    memset(a, 0, sizeof *a);
    a->fvc_mute = a->audio_mute = -1;
    command_processor::sptr b = command_processor::make();
    recorder rec;
    const char *ports[] = {"focc_words", "debug_output", "fvc_words", "audio_mute", "fvc_mute"};
    for (int i = 0; i < 5; ++i) rec.listen(*b, ports[i]);
    b->dispatch_msg("commands", pmt::cons(pmt::make_dict(), pmt::init_u8vector(strlen(cmd), (const uint8_t *)cmd)));
    for (size_t i = 0; i < rec.log.size(); ++i) {
        const recorded &r = rec.log[i];
        if (r.port == "focc_words") {
            a->focc_stream = pmt::to_long(pmt::tuple_ref(r.msg, 0));
            a->n_focc = (int32_t)pmt::to_long(pmt::tuple_ref(r.msg, 1));
            for (int w = 0; w < a->n_focc && w < 2; ++w) memcpy(a->focc_words[w], pmt::blob_data(pmt::tuple_ref(r.msg, 2 + w)), 28);
        } else if (r.port == "fvc_words") {
            a->has_fvc = 1;
            memcpy(a->fvc_word, pmt::blob_data(pmt::tuple_ref(r.msg, 1)), 28);
        } else if (r.port == "fvc_mute") {
            a->fvc_mute = pmt::to_bool(r.msg);
        } else if (r.port == "audio_mute") {
            a->audio_mute = pmt::to_bool(r.msg);
        } else if (r.port == "debug_output" && a->n_debug < 2) {
            put_text(a->debug[a->n_debug++], sizeof a->debug[0], pdu_text(r.msg));
        }
    }
}

/* ---- word builders and helpers (lib/amps_packet.cc, lib/amps_packet.h, lib/utils.cc, lib/focc_impl.cc) */
REF_API void ref_focc_word1(uint8_t w[28], int multiword, unsigned dcc, uint64_t min1) { focc_word1(w, multiword != 0, (unsigned char)dcc, min1); }
REF_API void ref_focc_word2_general(uint8_t w[28], uint64_t min2, unsigned msg_type, unsigned ordq, unsigned order) {
    focc_word2_general(w, min2, (unsigned char)msg_type, (unsigned char)ordq, (unsigned char)order);
}
REF_API void ref_fvc_word1_general(uint8_t w[28], unsigned pscc, unsigned msg_type, unsigned ordq, unsigned order) {
    fvc_word1_general(w, (unsigned char)pscc, (unsigned char)msg_type, (unsigned char)ordq, (unsigned char)order);
}
REF_API void ref_focc_word2_voice_channel(uint8_t w[28], unsigned scc, uint64_t min2, unsigned vmac, unsigned chan) {
    focc_word2_voice_channel(w, (unsigned char)scc, min2, (unsigned char)vmac, (unsigned short)chan);
}
REF_API void ref_expandbits(uint8_t *out, size_t nbits, uint64_t val) { expandbits(out, nbits, val); }
REF_API size_t ref_manchester_decode(const uint8_t *src, uint8_t *dst, size_t dstlen) { return manchester_decode_binbuf(src, dst, dstlen); }
REF_API void ref_extract_min_3(uint64_t val, char out3[3]) { std::string s = extract_min_3(val); memcpy(out3, s.data(), 3); }
REF_API uint64_t ref_compute_min_3(char d1, char d2, char d3) { return compute_min_3(d1, d2, d3); }
REF_API int ref_parse_min(const char *min, uint64_t *min1, uint64_t *min2) {
    u_int64_t a = 0, b = 0;
    bool ok = parse_min(min, a, b);
    *min1 = a; *min2 = b;
    return ok;
}
REF_API void ref_calc_min(uint64_t min1, uint64_t min2, char out11[11]) { put_text(out11, 11, calc_min((u_int64_t)min1, (u_int64_t)min2)); }
/* BCH(40,28) exactly as fvc_impl::fvc_bch / focc_impl::focc_bch do it: 23 pad zeros, itpp encode, drop 23 */
REF_API void ref_bch_encode_40_28(const uint8_t info28[28], uint8_t out40[40]) {
    itpp::BCH bch(63, 2, true);
    itpp::bvec padded(51);
    for (int i = 0; i < 28; ++i) padded[23 + i] = info28[i];
    itpp::bvec enc = bch.encode(padded);
    for (int i = 0; i < 40; ++i) out40[i] = (uint8_t)(int)enc[23 + i];
}
/* recc_decode_impl::recc_bch_decode (lib/recc_decode_impl.cc:53-79): validity of one 48-bit repeat */
REF_API int ref_bch_decode_48(const uint8_t in48[48], uint8_t out48[48]) {
    stdout_capture q;
    recc_decode_impl dec;
    unsigned char tmp[48];
    bool ok = dec.recc_bch_decode(in48, tmp);
    if (out48) memcpy(out48, tmp, 36);     /* only 36 corrected information bits exist (final = decoded(15, 50)) */
    return ok;
}

}  // extern "C"
