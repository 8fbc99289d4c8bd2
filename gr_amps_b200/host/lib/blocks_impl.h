// blocks_impl.h -- private implementation classes of the gr::amps blocks over libamps_b200's C ABI.
// Mirrors lib/{focc,fvc,recc,recc_decode}_impl.h of the reference (same members where they still exist).
#pragma once
#include <amps/focc.h>
#include <amps/fvc.h>
#include <amps/recc.h>
#include <amps/recc_decode.h>
#include <amps/recc_iq.h>
#include <amps/forward_iq.h>
#include <amps/command_processor.h>
#include <vector>
#include <amps_b200.h>

#include <string>

namespace gr { namespace amps {

enum focc_streams { STREAM_A = 1, STREAM_B = 2, STREAM_BOTH = 3 };     // lib/amps_packet.h:30-34
#define GLOBAL_SID 16
#define GLOBAL_DCC_SHORT 0
#define GLOBAL_SCC 1

// One C-ABI handle = one thread at a time.  GNU Radio runs a block's handlers and its work() on the block's own thread; the
// mutexes make the blocks safe under the stronger condition the reference also guards against (lib/focc_impl.cc:567,573):
// a handler entered from another thread while work() runs.
class focc_impl : public focc {
    amps_focc *d_h;
    boost::mutex d_mutex;
public:
    focc_impl(unsigned long symrate, bool aggressive_registration);
    ~focc_impl();
    void focc_words_message(pmt::pmt_t msg);
    int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);
};

class fvc_impl : public fvc {
    amps_fvc *d_h;
    boost::mutex d_mutex;
public:
    explicit fvc_impl(unsigned long symrate);
    ~fvc_impl();
    void fvc_words_message(pmt::pmt_t msg);
    int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);
};

class recc_impl : public recc {
    amps_recc *d_h;
    static void on_blob(const uint8_t *blob, void *self);
public:
    recc_impl();
    ~recc_impl();
    int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);
};

class recc_iq_impl : public recc_iq {
    amps_recc_iq *d_h;
    static void on_burst(const amps_burst *b, void *self);
public:
    bool d_sc16;
    recc_iq_impl(double samp_rate, double center_freq, int device, bool mm_timing, bool sc16);
    ~recc_iq_impl();
    int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);
};

class recc_decode_impl : public recc_decode {
    amps_recc_decode *d_h;
public:
    recc_decode_impl();
    ~recc_decode_impl();
    void bursts_message(pmt::pmt_t msg);
    void handle_origination(const amps_recc_words &w);
    void handle_response(const amps_recc_words &w);
    void handle_registration(const amps_recc_words &w);
};

class command_processor_impl : public command_processor {            // host-only: text commands -> control words
    void debug_msg(const char *msg);
    void handle_page(const std::string numstr);
public:
    command_processor_impl();
    void commands_message(pmt::pmt_t msg);
};

class forward_iq_impl : public forward_iq {
    amps_focc *d_focc;
    amps_fvc *d_fvc;
    amps_fwd *d_fwd;
    bool d_fvc_mute;                       // the reference graph starts with the FVC leg muted (grc/ampsbs.grc:1601)
    boost::mutex d_mutex;
    std::vector<uint8_t> d_bits[3];
public:
    forward_iq_impl(bool aggressive_registration, int device);
    ~forward_iq_impl();
    void focc_words_message(pmt::pmt_t msg);
    void fvc_words_message(pmt::pmt_t msg);
    void fvc_mute_message(pmt::pmt_t msg);
    int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);
};

}}  // namespace gr::amps
