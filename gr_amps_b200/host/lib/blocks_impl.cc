// blocks_impl.cc -- gr::amps::{focc,fvc,recc,recc_iq,recc_decode} on top of libamps_b200.
// Same names, make() signatures, stream signatures, message ports and work() return conventions as the
// reference (lib/*_impl.cc); the bodies are calls into the C ABI.  Errors follow the reference's style:
// construction failures throw std::runtime_error (a flowgraph cannot start), run-time failures are logged.
#include "blocks_impl.h"
#include "../../csrc/proto.h"

#include <algorithm>
#include <cstdio>
#include <cctype>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace gr { namespace amps {

static void must(int status, const char *what) {
    if (status != AMPS_OK) throw std::runtime_error(std::string(what) + ": " + amps_b200_last_error());
}
static void warn(int status, const char *what) {
    if (status != AMPS_OK) std::fprintf(stderr, "gr-amps(b200) %s: %s\n", what, amps_b200_last_error());
}

// (first, first + 1, ...) = nwords blobs of 28 bytes each, copied into one buffer.  The reference asserts nwords == len - 2
// (lib/focc_impl.cc:527, compiled out in its Release build); a malformed tuple is dropped with a warning here.
static bool words_from_tuple(pmt::pmt_t msg, size_t first, long nwords, size_t max_extra, std::vector<uint8_t> &w, const char *what) {
    const size_t len = pmt::length(msg);
    if (nwords < 0 || nwords > 4096 || first + (size_t)nwords > len || len > first + (size_t)nwords + max_extra) {
        std::fprintf(stderr, "gr-amps(b200) %s: malformed message (%ld words announced, %zu elements): dropped\n", what, nwords, len);
        return false;
    }
    w.resize((size_t)28 * (size_t)nwords);
    for (long i = 0; i < nwords; i++) {
        pmt::pmt_t blob = pmt::tuple_ref(msg, first + (size_t)i);
        if (!pmt::is_blob(blob) || pmt::blob_length(blob) != 28) {
            std::fprintf(stderr, "gr-amps(b200) %s: word %ld is not a 28-byte blob: message dropped\n", what, i);
            return false;
        }
        std::memcpy(&w[(size_t)28 * (size_t)i], pmt::blob_data(blob), 28);
    }
    return true;
}

// ------------------------------------------------------------------ focc (lib/focc_impl.cc)
focc::sptr focc::make(unsigned long symrate, bool aggressive_registration) {
    return gnuradio::get_initial_sptr(new focc_impl(symrate, aggressive_registration));
}
focc_impl::focc_impl(unsigned long symrate, bool aggressive_registration)
    : gr::sync_block("focc", gr::io_signature::make(0, 0, 0), gr::io_signature::make(1, 1, sizeof(unsigned char))), d_h(NULL) {
    must(amps_focc_create(symrate, aggressive_registration ? 1 : 0, 0, &d_h), "focc");
    message_port_register_in(pmt::mp("focc_words"));                                   // lib/focc_impl.cc:127-130
    set_msg_handler(pmt::mp("focc_words"), boost::bind(&focc_impl::focc_words_message, this, _1));
}
focc_impl::~focc_impl() { amps_focc_destroy(d_h); }
void focc_impl::focc_words_message(pmt::pmt_t msg) {                                   // lib/focc_impl.cc:521-563
    if (!pmt::is_tuple(msg) || pmt::length(msg) < 3) return;
    const long stream = pmt::to_long(pmt::tuple_ref(msg, 0));
    const long nwords = pmt::to_long(pmt::tuple_ref(msg, 1));
    std::vector<uint8_t> w;
    if (!words_from_tuple(msg, 2, nwords, 0, w, "focc_words")) return;
    boost::mutex::scoped_lock lock(d_mutex);                                            // lib/focc_impl.cc:567
    warn(amps_focc_push_words(d_h, stream, w.data(), nwords), "focc_words");
}
int focc_impl::work(int noutput_items, gr_vector_const_void_star &, gr_vector_void_star &output_items) {
    int produced = 0;
    boost::mutex::scoped_lock lock(d_mutex);                                            // lib/focc_impl.cc:573
    if (amps_focc_work(d_h, static_cast<uint8_t *>(output_items[0]), noutput_items, &produced) != AMPS_OK) {
        warn(AMPS_E_CUDA, "focc work");
        return WORK_DONE;                                                               // a dead device ends the flowgraph; it does not emit garbage
    }
    return produced;      // <= one burst, possibly 0, -1 (WORK_DONE) when noutput_items < 1 (lib/focc_impl.cc:590-593,630-632)
}

// ------------------------------------------------------------------ fvc (lib/fvc_impl.cc)
fvc::sptr fvc::make(unsigned long symrate) { return gnuradio::get_initial_sptr(new fvc_impl(symrate)); }
fvc_impl::fvc_impl(unsigned long symrate)
    : gr::sync_block("fvc", gr::io_signature::make(0, 0, 0), gr::io_signature::make(1, 1, sizeof(unsigned char))), d_h(NULL) {
    must(amps_fvc_create(symrate, 0, &d_h), "fvc");
    message_port_register_in(pmt::mp("fvc_words"));                                    // lib/fvc_impl.cc:62-66
    set_msg_handler(pmt::mp("fvc_words"), boost::bind(&fvc_impl::fvc_words_message, this, _1));
    message_port_register_out(pmt::mp("command_out"));
}
fvc_impl::~fvc_impl() { amps_fvc_destroy(d_h); }
void fvc_impl::fvc_words_message(pmt::pmt_t msg) {                                     // lib/fvc_impl.cc:109-143
    if (!pmt::is_tuple(msg) || pmt::length(msg) < 2) return;
    const size_t len = pmt::length(msg);
    const long nwords = pmt::to_long(pmt::tuple_ref(msg, 0));
    std::vector<uint8_t> w;
    if (!words_from_tuple(msg, 1, nwords, 1, w, "fvc_words")) return;                   // (+ the optional trailing uint64 timer)
    const bool has_timer = len > (size_t)(1 + nwords);
    const uint64_t timer = has_timer ? pmt::to_uint64(pmt::tuple_ref(msg, 1 + (size_t)nwords)) : 0;
    boost::mutex::scoped_lock lock(d_mutex);
    warn(amps_fvc_push_words(d_h, w.data(), nwords, has_timer ? 1 : 0, timer), "fvc_words");
}
int fvc_impl::work(int noutput_items, gr_vector_const_void_star &, gr_vector_void_star &output_items) {
    int produced = 0, off = 0;
    {
        boost::mutex::scoped_lock lock(d_mutex);
        if (amps_fvc_work(d_h, static_cast<uint8_t *>(output_items[0]), noutput_items, &produced, &off) != AMPS_OK) {
            warn(AMPS_E_CUDA, "fvc work");
            return WORK_DONE;
        }
    }
    if (off) {                                                                          // lib/fvc_impl.cc:163-171
        const char *m = "fvc off";
        message_port_pub(pmt::mp("command_out"), pmt::cons(pmt::make_dict(), pmt::init_u8vector(std::strlen(m), (const uint8_t *)m)));
    }
    return produced;
}

// ------------------------------------------------------------------ recc (lib/recc_impl.cc)
recc::sptr recc::make() { return gnuradio::get_initial_sptr(new recc_impl()); }
recc_impl::recc_impl()
    : gr::sync_block("recc", gr::io_signature::make(1, 1, sizeof(unsigned char)), gr::io_signature::make(0, 0, 0)), d_h(NULL) {
    must(amps_recc_create(0, &d_h), "recc");
    message_port_register_out(pmt::mp("bursts"));                                      // lib/recc_impl.cc:82
}
recc_impl::~recc_impl() { amps_recc_destroy(d_h); }
void recc_impl::on_blob(const uint8_t *blob, void *self) {
    static_cast<recc_impl *>(self)->message_port_pub(pmt::mp("bursts"), pmt::mp(blob, AMPS_RECC_CAPTURE_SYMS));   // :126
}
int recc_impl::work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &) {
    if (noutput_items < 1) return 0;                                                   // :98-101
    if (amps_recc_work(d_h, static_cast<const uint8_t *>(input_items[0]), noutput_items, &recc_impl::on_blob, this) != AMPS_OK) {
        warn(AMPS_E_CUDA, "recc work");
        return WORK_DONE;                                                               // not consumed: nothing is silently dropped
    }
    consume_each(noutput_items);                                                       // :113
    return 0;                                                                          // :144
}

// ------------------------------------------------------------------ recc_iq (new sibling block)
static const uint32_t kMaxItemsPerWork = 1u << 22;
recc_iq::sptr recc_iq::make(double samp_rate, double center_freq, int device, bool mm_timing, bool sc16) {
    return gnuradio::get_initial_sptr(new recc_iq_impl(samp_rate, center_freq, device, mm_timing, sc16));
}
recc_iq_impl::recc_iq_impl(double samp_rate, double center_freq, int device, bool mm_timing, bool sc16)
    : gr::sync_block("recc_iq", gr::io_signature::make(1, 1, sc16 ? 2 * sizeof(short) : sizeof(std::complex<float>)), gr::io_signature::make(0, 0, 0)),
      d_h(NULL), d_sc16(sc16) {
    amps_recc_iq_params p;
    std::memset(&p, 0, sizeof p);
    p.samp_rate = samp_rate; p.center_freq = center_freq; p.device = device;
    p.max_samples = kMaxItemsPerWork;
    p.flags = (mm_timing ? AMPS_RX_TIMING_MM : 0u) | (sc16 ? AMPS_RX_INPUT_SC16 : 0u);
    must(amps_recc_iq_create(&p, &d_h), "recc_iq");
    set_max_noutput_items((int)kMaxItemsPerWork);  // the handle's buffers are sized for it: the scheduler never hands the block more at once
    message_port_register_out(pmt::mp("bursts"));
}
recc_iq_impl::~recc_iq_impl() { amps_recc_iq_destroy(d_h); }
void recc_iq_impl::on_burst(const amps_burst *b, void *self) {
    static_cast<recc_iq_impl *>(self)->message_port_pub(pmt::mp("bursts"), pmt::mp(b->symbols, AMPS_RECC_CAPTURE_SYMS));
}
int recc_iq_impl::work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &) {
    if (noutput_items < 1) return 0;
    // gr_complex is std::complex<float>: interleaved re, im -- exactly what the ABI takes.  A scheduler that ignores
    // set_max_noutput_items is served in slices.
    size_t done = 0;
    while (done < (size_t)noutput_items) {
        const size_t n = std::min((size_t)noutput_items - done, (size_t)kMaxItemsPerWork);
        const int st = d_sc16
            ? amps_recc_iq_work_sc16(d_h, static_cast<const int16_t *>(input_items[0]) + 2 * done, n, &recc_iq_impl::on_burst, this)
            : amps_recc_iq_work(d_h, static_cast<const float *>(input_items[0]) + 2 * done, n, &recc_iq_impl::on_burst, this);
        if (st == AMPS_E_OVERFLOW) {
            warn(st, "recc_iq work (candidate list overflow: bursts may have been missed; the stream goes on)");
        } else if (st != AMPS_OK) {
            warn(st, "recc_iq work");
            // samples must not disappear silently: report what was taken so far and end the flowgraph
            return done ? (int)done : WORK_DONE;
        }
        done += n;
    }
    return noutput_items;
}

// ------------------------------------------------------------------ recc_decode (lib/recc_decode_impl.cc)
recc_decode::sptr recc_decode::make() { return gnuradio::get_initial_sptr(new recc_decode_impl()); }
recc_decode_impl::recc_decode_impl()
    : gr::block("recc_decode", gr::io_signature::make(0, 0, 0), gr::io_signature::make(0, 0, 0)), d_h(NULL) {
    must(amps_recc_decode_create(0, &d_h), "recc_decode");
    message_port_register_in(pmt::mp("bursts"));                                       // lib/recc_decode_impl.cc:38-46
    set_msg_handler(pmt::mp("bursts"), boost::bind(&recc_decode_impl::bursts_message, this, _1));
    message_port_register_out(pmt::mp("focc_words"));
    message_port_register_out(pmt::mp("fvc_words"));
    message_port_register_out(pmt::mp("audio_mute"));
    message_port_register_out(pmt::mp("fvc_mute"));
    message_port_register_out(pmt::mp("command_out"));
}
recc_decode_impl::~recc_decode_impl() { amps_recc_decode_destroy(d_h); }

void recc_decode_impl::bursts_message(pmt::pmt_t msg) {                                // lib/recc_decode_impl.cc:81-169
    if (!pmt::is_blob(msg) || pmt::blob_length(msg) < AMPS_RECC_CAPTURE_SYMS) return;
    amps_recc_words w;
    if (amps_recc_decode_burst(d_h, static_cast<const uint8_t *>(pmt::blob_data(msg)), &w) != AMPS_OK) {
        warn(AMPS_E_CUDA, "recc_decode");
        return;
    }
    switch (w.kind) {
        case AMPS_MSG_INVALID_A: return;                                               // :108-111
        case AMPS_MSG_E0_DROPPED: return;                                              // :113-116
        case AMPS_MSG_PAGE_RESPONSE: handle_response(w); return;                       // :121
        case AMPS_MSG_REGISTRATION: handle_registration(w); return;                    // :123-138
        case AMPS_MSG_ORIGINATION: handle_origination(w); return;                      // :139-165
        default: return;                                                               // unknown / bad NAWC: logged and dropped
    }
}

static pmt::pmt_t word_blob(const ::amps::Word28 &w) { return pmt::mp(w.data(), 28); }

void recc_decode_impl::handle_registration(const amps_recc_words &w) {                 // :181-190: order confirmation (audit, order 7)
    const ::amps::Word28 w1 = ::amps::focc_word1(true, GLOBAL_DCC_SHORT, w.MIN1);
    const ::amps::Word28 w2 = ::amps::focc_word2_general(w.MIN2, 0, 0, 7);
    message_port_pub(pmt::mp("focc_words"), pmt::make_tuple(pmt::from_long(STREAM_BOTH), pmt::from_long(2), word_blob(w1), word_blob(w2)));
}

void recc_decode_impl::handle_response(const amps_recc_words &w) {                     // :195-220: page response -> voice channel 355 + alert
    const ::amps::Word28 w1 = ::amps::focc_word1(true, GLOBAL_DCC_SHORT, w.MIN1);
    const ::amps::Word28 w2 = ::amps::focc_word2_voice_channel(GLOBAL_SCC, w.MIN2, 0, 355);
    message_port_pub(pmt::mp("focc_words"), pmt::make_tuple(pmt::from_long(STREAM_BOTH), pmt::from_long(2), word_blob(w1), word_blob(w2)));
    const ::amps::Word28 fv = ::amps::fvc_word1_general(GLOBAL_SCC, 0, 0, 1);
    message_port_pub(pmt::mp("fvc_words"), pmt::make_tuple(pmt::from_long(1), word_blob(fv), pmt::from_uint64(35)));
    message_port_pub(pmt::mp("fvc_mute"), pmt::from_bool(false));
    message_port_pub(pmt::mp("audio_mute"), pmt::from_bool(true));
}

void recc_decode_impl::handle_origination(const amps_recc_words &w) {                  // :234-272: initial voice designation, channel 356
    const ::amps::Word28 w1 = ::amps::focc_word1(true, GLOBAL_DCC_SHORT, w.MIN1);
    const ::amps::Word28 w2 = (w.dialed[0] == '0') ? ::amps::focc_word2_general(w.MIN2, 0, 0, 9)
                                                    : ::amps::focc_word2_voice_channel(GLOBAL_SCC, w.MIN2, 0, 356);
    message_port_pub(pmt::mp("focc_words"), pmt::make_tuple(pmt::from_long(STREAM_BOTH), pmt::from_long(2), word_blob(w1), word_blob(w2)));
    message_port_pub(pmt::mp("fvc_mute"), pmt::from_bool(true));
    message_port_pub(pmt::mp("audio_mute"), pmt::from_bool(false));
    const std::string m = std::string("page ") + w.dialed;
    message_port_pub(pmt::mp("command_out"), pmt::cons(pmt::make_dict(), pmt::init_u8vector(m.size(), (const uint8_t *)m.data())));
}

// ------------------------------------------------------------------ command_processor (lib/command_processor_impl.cc)
command_processor::sptr command_processor::make() { return gnuradio::get_initial_sptr(new command_processor_impl()); }
command_processor_impl::command_processor_impl()
    : gr::block("command_processor", gr::io_signature::make(0, 0, 0), gr::io_signature::make(0, 0, 0)) {
    message_port_register_in(pmt::mp("commands"));                                     // :41-49
    set_msg_handler(pmt::mp("commands"), boost::bind(&command_processor_impl::commands_message, this, _1));
    message_port_register_out(pmt::mp("focc_words"));
    message_port_register_out(pmt::mp("debug_output"));
    message_port_register_out(pmt::mp("fvc_words"));
    message_port_register_out(pmt::mp("audio_mute"));
    message_port_register_out(pmt::mp("fvc_mute"));
}

void command_processor_impl::debug_msg(const char *msg) {                              // :52-56
    message_port_pub(pmt::mp("debug_output"), pmt::cons(pmt::make_dict(), pmt::init_u8vector(std::strlen(msg), (const uint8_t *)msg)));
}

void command_processor_impl::handle_page(const std::string numstr) {                   // :58-82: page = Word 1 + Word 2, order 0
    if (numstr.empty()) { debug_msg("missing MIN in page command\n"); return; }
    debug_msg("paging!\n");
    uint64_t min1, min2;
    if (!::amps::parse_min(numstr, min1, min2)) { debug_msg("invalid MIN entered"); return; }
    const ::amps::Word28 w1 = ::amps::focc_word1(true, GLOBAL_DCC_SHORT, min1);
    const ::amps::Word28 w2 = ::amps::focc_word2_general(min2, 0, 0, 0);
    message_port_pub(pmt::mp("focc_words"), pmt::make_tuple(pmt::from_long(STREAM_BOTH), pmt::from_long(2), word_blob(w1), word_blob(w2)));
}

static bool has_prefix(const std::string &s, const char *p, bool any_case) {
    for (size_t i = 0; p[i]; i++) {
        if (i >= s.size()) return false;
        const unsigned char a = (unsigned char)s[i], b = (unsigned char)p[i];
        if (any_case ? std::tolower(a) != std::tolower(b) : a != b) return false;
    }
    return true;
}

void command_processor_impl::commands_message(pmt::pmt_t msg) {                        // :84-113: PDU (dict . u8vector of text)
    if (!pmt::is_pair(msg) || !pmt::is_u8vector(pmt::cdr(msg))) return;
    size_t n = 0;
    const uint8_t *chars = pmt::u8vector_elements(pmt::cdr(msg), n);
    std::string cmd(reinterpret_cast<const char *>(chars), n);
    cmd = cmd.substr(0, cmd.find('\0'));                                               // the reference goes through a C string
    if (has_prefix(cmd, "fvc off", false)) {
        message_port_pub(pmt::mp("fvc_mute"), pmt::from_bool(true));
        message_port_pub(pmt::mp("audio_mute"), pmt::from_bool(false));
        debug_msg("turning FVC data OFF; audio ON\n");
    } else if (has_prefix(cmd, "fvc on", false)) {
        message_port_pub(pmt::mp("fvc_mute"), pmt::from_bool(false));
        message_port_pub(pmt::mp("audio_mute"), pmt::from_bool(true));
        debug_msg("turning FVC data ON; audio OFF\n");
    } else if (has_prefix(cmd, "fvc alert", false)) {
        const ::amps::Word28 fv = ::amps::fvc_word1_general(GLOBAL_SCC, 0, 0, 1);
        message_port_pub(pmt::mp("fvc_words"), pmt::make_tuple(pmt::from_long(1), word_blob(fv)));
    } else if (has_prefix(cmd, "page ", true)) {
        size_t b = 5, e = cmd.size();
        while (b < e && std::isspace((unsigned char)cmd[b])) b++;
        while (e > b && std::isspace((unsigned char)cmd[e - 1])) e--;
        handle_page(cmd.substr(b, e - b));
    } else {
        debug_msg("invalid command\n");
    }
}

// ------------------------------------------------------------------ forward_iq (new composite source block)
static const int kSamplesPerBit = 1000;            // 10 MS/s / 10 kbit/s
static const size_t kMaxBitsPerWork = 4096;

forward_iq::sptr forward_iq::make(bool aggressive_registration, int device) {
    return gnuradio::get_initial_sptr(new forward_iq_impl(aggressive_registration, device));
}
forward_iq_impl::forward_iq_impl(bool aggressive_registration, int device)
    : gr::sync_block("forward_iq", gr::io_signature::make(0, 0, 0), gr::io_signature::make(1, 1, sizeof(std::complex<float>))),
      d_focc(NULL), d_fvc(NULL), d_fwd(NULL), d_fvc_mute(true) {
    must(amps_focc_create(100000, aggressive_registration ? 1 : 0, device, &d_focc), "forward_iq/focc");
    must(amps_fvc_create(100000, device, &d_fvc), "forward_iq/fvc");
    amps_fwd_params p;
    std::memset(&p, 0, sizeof p);
    p.samp_rate = 10e6; p.symrate = 100e3; p.max_deviation = 8000; p.device = device; p.ncarriers = 3;
    p.carrier_freq[0] = 0; p.carrier_freq[1] = 60e3; p.carrier_freq[2] = 90e3;          // grc/ampsbs.grc:841,904
    p.lpf_transition[0] = 5e3; p.lpf_transition[1] = 3e3; p.lpf_transition[2] = 3e3;    // :2227, :2172
    p.out_scale = 0.5f;                                                                  // :1367
    p.max_samples = (uint32_t)(kMaxBitsPerWork * kSamplesPerBit);
    must(amps_fwd_create(&p, &d_fwd), "forward_iq/fwd");
    set_output_multiple(kSamplesPerBit);            // whole data bits per call
    message_port_register_in(pmt::mp("focc_words"));
    set_msg_handler(pmt::mp("focc_words"), boost::bind(&forward_iq_impl::focc_words_message, this, _1));
    message_port_register_in(pmt::mp("fvc_words"));
    set_msg_handler(pmt::mp("fvc_words"), boost::bind(&forward_iq_impl::fvc_words_message, this, _1));
    message_port_register_in(pmt::mp("fvc_mute"));
    set_msg_handler(pmt::mp("fvc_mute"), boost::bind(&forward_iq_impl::fvc_mute_message, this, _1));
    message_port_register_out(pmt::mp("command_out"));
}
forward_iq_impl::~forward_iq_impl() { amps_fwd_destroy(d_fwd); amps_fvc_destroy(d_fvc); amps_focc_destroy(d_focc); }

void forward_iq_impl::focc_words_message(pmt::pmt_t msg) {
    if (!pmt::is_tuple(msg) || pmt::length(msg) < 3) return;
    const long stream = pmt::to_long(pmt::tuple_ref(msg, 0)), nwords = pmt::to_long(pmt::tuple_ref(msg, 1));
    std::vector<uint8_t> w;
    if (!words_from_tuple(msg, 2, nwords, 0, w, "forward_iq focc_words")) return;
    boost::mutex::scoped_lock lock(d_mutex);
    warn(amps_focc_push_words(d_focc, stream, w.data(), nwords), "forward_iq focc_words");
}
void forward_iq_impl::fvc_words_message(pmt::pmt_t msg) {
    if (!pmt::is_tuple(msg) || pmt::length(msg) < 2) return;
    const size_t len = pmt::length(msg);
    const long nwords = pmt::to_long(pmt::tuple_ref(msg, 0));
    std::vector<uint8_t> w;
    if (!words_from_tuple(msg, 1, nwords, 1, w, "forward_iq fvc_words")) return;
    const bool has_timer = len > (size_t)(1 + nwords);
    boost::mutex::scoped_lock lock(d_mutex);
    warn(amps_fvc_push_words(d_fvc, w.data(), nwords, has_timer ? 1 : 0, has_timer ? pmt::to_uint64(pmt::tuple_ref(msg, 1 + (size_t)nwords)) : 0),
         "forward_iq fvc_words");
}
void forward_iq_impl::fvc_mute_message(pmt::pmt_t msg) {
    boost::mutex::scoped_lock lock(d_mutex);
    d_fvc_mute = pmt::to_bool(msg);
}

int forward_iq_impl::work(int noutput_items, gr_vector_const_void_star &, gr_vector_void_star &output_items) {
    size_t nbits = (size_t)noutput_items / kSamplesPerBit;
    if (nbits > kMaxBitsPerWork) nbits = kMaxBitsPerWork;
    if (nbits == 0) return 0;
    boost::mutex::scoped_lock lock(d_mutex);
    for (int c = 0; c < 3; c++) d_bits[c].assign(nbits, 0xFF);
    warn(amps_focc_generate_bits(d_focc, d_bits[0].data(), nbits), "forward_iq focc bits");
    // the FVC source runs whether or not its leg is muted (as the fvc block does behind mute_xx)
    size_t got = 0;
    while (got < nbits) {
        int produced = 0, off = 0;
        if (amps_fvc_work_bits(d_fvc, d_bits[1].data() + got, (int)(nbits - got), &produced, &off) != AMPS_OK || produced <= 0) break;
        if (off) {
            const char *m = "fvc off";
            message_port_pub(pmt::mp("command_out"), pmt::cons(pmt::make_dict(), pmt::init_u8vector(std::strlen(m), (const uint8_t *)m)));
        }
        got += (size_t)produced;
    }
    if (d_fvc_mute) std::fill(d_bits[1].begin(), d_bits[1].end(), (uint8_t)0xFF);
    const uint8_t *bits[3] = {d_bits[0].data(), d_bits[1].data(), d_bits[2].data()};
    if (amps_fwd_work_bits(d_fwd, bits, nbits, static_cast<float *>(output_items[0])) != AMPS_OK) {
        warn(AMPS_E_CUDA, "forward_iq work");
        return WORK_DONE;
    }
    return (int)(nbits * kSamplesPerBit);
}

}}  // namespace gr::amps
