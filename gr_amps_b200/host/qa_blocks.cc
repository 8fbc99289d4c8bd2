// qa_blocks.cc -- C++ QA driver for the gr::amps blocks built on libamps_b200 (the reference's own
// lib/qa_amps.cc is an empty CppUnit suite; apps/testalloc.cc is its only harness and is the model here).
// It wires the blocks exactly as grc/ampsbs.grc:4404-4470 does (msg_connect) and drives work() the way the
// GNU Radio scheduler would.  Results go to files / JSON lines that tests/test_host_gpu.py compares with the oracle.
//
//   qa_blocks focc <symrate> <aggr 0|1> <total_bytes> <seed> <out.bin>
//   qa_blocks loop <iq.bin> <nsamples> <chunk> <focc_bytes> <out_prefix> [mm]
//   qa_blocks fwd <nsym> <out.bin> [voice]
//   qa_blocks cmd <text> [<text> ...]          (host only: command_processor, no GPU needed)
//   qa_blocks threads <total_bytes> <nmsgs> <out.bin>   (focc_words posted from a second thread while work() runs)
//   qa_blocks badmsg                           (malformed focc_words / fvc_words tuples are dropped, not crashed on)
//   qa_blocks batch <iq.bin> <nsamples> <K> <chunk>   (C ABI: K carriers on one uploaded buffer, amps_recc_iq_batch_work_shared)
#include <amps/focc.h>
#include <amps/fvc.h>
#include <amps/recc.h>
#include <amps/recc_decode.h>
#include <amps/recc_iq.h>
#include <amps/forward_iq.h>
#include <amps/command_processor.h>
#include <amps_b200.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

using namespace gr::amps;

static std::string bits28(const void *p) {
    std::string s;
    for (int i = 0; i < 28; i++) s += static_cast<const char *>(p)[i] ? '1' : '0';
    return s;
}

// message sink that records everything recc_decode / fvc publish, as JSON lines
class probe : public gr::block {
public:
    std::vector<std::string> lines;
    probe() : gr::block("probe", gr::io_signature::make(0, 0, 0), gr::io_signature::make(0, 0, 0)) {
        const char *ports[] = {"focc_words", "fvc_words", "audio_mute", "fvc_mute", "command_out", "bursts", "debug_output"};
        for (size_t i = 0; i < sizeof(ports) / sizeof(ports[0]); i++) {
            const std::string port = ports[i];
            message_port_register_in(pmt::mp(port));
            set_msg_handler(pmt::mp(port), [this, port](pmt::pmt_t m) { this->on(port, m); });
        }
    }
    void on(const std::string &port, pmt::pmt_t m) {
        char buf[256];
        std::string js = "{\"port\": \"" + port + "\"";
        if (port == "focc_words") {
            const long n = pmt::to_long(pmt::tuple_ref(m, 1));
            std::snprintf(buf, sizeof buf, ", \"stream\": %ld, \"n\": %ld, \"words\": [", pmt::to_long(pmt::tuple_ref(m, 0)), n);
            js += buf;
            for (long i = 0; i < n; i++) js += std::string(i ? ", " : "") + "\"" + bits28(pmt::blob_data(pmt::tuple_ref(m, 2 + (size_t)i))) + "\"";
            js += "]";
        } else if (port == "fvc_words") {
            const long n = pmt::to_long(pmt::tuple_ref(m, 0));
            js += ", \"words\": [";
            for (long i = 0; i < n; i++) js += std::string(i ? ", " : "") + "\"" + bits28(pmt::blob_data(pmt::tuple_ref(m, 1 + (size_t)i))) + "\"";
            js += "]";
            if (pmt::length(m) > (size_t)(1 + n)) {
                std::snprintf(buf, sizeof buf, ", \"timer\": %llu", (unsigned long long)pmt::to_uint64(pmt::tuple_ref(m, 1 + (size_t)n)));
                js += buf;
            }
        } else if (port == "audio_mute" || port == "fvc_mute") {
            js += std::string(", \"value\": ") + (pmt::to_bool(m) ? "true" : "false");
        } else if (port == "command_out" || port == "debug_output") {
            size_t n = 0;
            const uint8_t *d = pmt::u8vector_elements(pmt::cdr(m), n);
            std::string text;
            for (size_t i = 0; i < n; i++) text += (d[i] == '\n') ? std::string("\\n") : std::string(1, (char)d[i]);
            js += ", \"text\": \"" + text + "\"";
        } else if (port == "bursts") {
            std::snprintf(buf, sizeof buf, ", \"len\": %zu", pmt::blob_length(m));
            js += buf;
        }
        lines.push_back(js + "}");
    }
};

static int run_focc(int argc, char **argv) {
    if (argc < 7) return 2;
    const unsigned long symrate = std::strtoul(argv[2], NULL, 10);
    const bool aggr = std::atoi(argv[3]) != 0;
    const size_t total = std::strtoull(argv[4], NULL, 10);
    unsigned long long lcg = std::strtoull(argv[5], NULL, 10);
    focc::sptr blk = focc::make(symrate, aggr);
    std::vector<unsigned char> out, buf(1 << 16);
    gr_vector_const_void_star in;
    gr_vector_void_star outs(1);
    while (out.size() < total) {
        lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
        int n = 1 + (int)((lcg >> 33) % 9000);
        if ((size_t)n > total - out.size()) n = (int)(total - out.size());
        outs[0] = buf.data();
        const int r = blk->work(n, in, outs);              // apps/testalloc.cc:57: retval may be short
        if (r < 0) return 3;
        out.insert(out.end(), buf.begin(), buf.begin() + r);
    }
    std::ofstream(argv[6], std::ios::binary).write(reinterpret_cast<const char *>(out.data()), (std::streamsize)out.size());
    std::printf("{\"bytes\": %zu}\n", out.size());
    return 0;
}

static int run_loop(int argc, char **argv) {
    if (argc < 7) return 2;
    const size_t nsamples = std::strtoull(argv[3], NULL, 10);
    const size_t chunk = std::strtoull(argv[4], NULL, 10);
    const size_t focc_bytes = std::strtoull(argv[5], NULL, 10);
    const std::string prefix = argv[6];
    std::vector<std::complex<float>> iq(nsamples);
    std::ifstream f(argv[2], std::ios::binary);
    f.read(reinterpret_cast<char *>(iq.data()), (std::streamsize)(nsamples * sizeof(std::complex<float>)));
    if (!f) { std::fprintf(stderr, "short read of %s\n", argv[2]); return 4; }

    // the message wiring of grc/ampsbs.grc:4404-4470 around the RECC/FOCC/FVC blocks
    const bool mm_timing = argc > 7 && !std::strcmp(argv[7], "mm");                   // M&M timing tail instead of the detector
    const bool sc16 = argc > 7 && !std::strcmp(argv[7], "sc16");                     // int16 I,Q input (x 1/32768 on the GPU)
    std::vector<int16_t> iq16;
    if (sc16) {
        iq16.resize(2 * nsamples);
        const float *f32 = reinterpret_cast<const float *>(iq.data());
        for (size_t i = 0; i < 2 * nsamples; i++) {
            float v = std::nearbyint(f32[i] * 32768.0f);
            iq16[i] = (int16_t)(v > 32767.f ? 32767.f : (v < -32768.f ? -32768.f : v));
        }
    }
    recc_iq::sptr rx = recc_iq::make(10e6, -160e3, 0, mm_timing, sc16);
    recc_decode::sptr dec = recc_decode::make();
    focc::sptr fo = focc::make(100000, false);
    fvc::sptr fv = fvc::make(100000);
    probe pr;
    gr::msg_connect(*rx, "bursts", *dec, "bursts");
    gr::msg_connect(*rx, "bursts", pr, "bursts");
    gr::msg_connect(*dec, "focc_words", *fo, "focc_words");
    gr::msg_connect(*dec, "fvc_words", *fv, "fvc_words");
    const char *ports[] = {"focc_words", "fvc_words", "audio_mute", "fvc_mute", "command_out"};
    for (size_t i = 0; i < 5; i++) gr::msg_connect(*dec, ports[i], pr, ports[i]);
    gr::msg_connect(*fv, "command_out", pr, "command_out");

    gr_vector_const_void_star in(1);
    gr_vector_void_star none;
    for (size_t pos = 0; pos < nsamples; pos += chunk) {
        const size_t n = nsamples - pos < chunk ? nsamples - pos : chunk;
        in[0] = sc16 ? static_cast<const void *>(&iq16[2 * pos]) : static_cast<const void *>(&iq[pos]);
        if (rx->work((int)n, in, none) != (int)n) return 5;
    }
    // then let the sources run, as the scheduler would
    std::vector<unsigned char> out, buf(1 << 16);
    gr_vector_const_void_star noin;
    gr_vector_void_star outs(1);
    while (out.size() < focc_bytes) {
        outs[0] = buf.data();
        int want = (int)(focc_bytes - out.size() < buf.size() ? focc_bytes - out.size() : buf.size());
        const int r = fo->work(want, noin, outs);
        if (r < 0) return 6;
        out.insert(out.end(), buf.begin(), buf.begin() + r);
    }
    std::ofstream((prefix + ".focc.bin").c_str(), std::ios::binary).write(reinterpret_cast<const char *>(out.data()), (std::streamsize)out.size());
    std::vector<unsigned char> vout;
    while (vout.size() < 30000) {
        outs[0] = buf.data();
        const int r = fv->work(4096, noin, outs);
        if (r < 0) return 7;
        vout.insert(vout.end(), buf.begin(), buf.begin() + r);
    }
    std::ofstream((prefix + ".fvc.bin").c_str(), std::ios::binary).write(reinterpret_cast<const char *>(vout.data()), (std::streamsize)vout.size());
    for (size_t i = 0; i < pr.lines.size(); i++) std::printf("%s\n", pr.lines[i].c_str());
    return 0;
}

// forward path straight through the C ABI: FOCC block bytes + two FVC alert trains -> amps_fwd_work -> file
static int run_fwd(int argc, char **argv) {
    if (argc < 4) return 2;
    const size_t nsym = std::strtoull(argv[2], NULL, 10);
    focc::sptr fo = focc::make(100000, false);
    std::vector<unsigned char> s0, buf(1 << 16);
    gr_vector_const_void_star noin;
    gr_vector_void_star outs(1);
    while (s0.size() < nsym) {
        outs[0] = buf.data();
        const int r = fo->work((int)(nsym - s0.size() < buf.size() ? nsym - s0.size() : buf.size()), noin, outs);
        if (r < 0) return 3;
        s0.insert(s0.end(), buf.begin(), buf.begin() + r);
    }
    std::vector<unsigned char> s1(nsym), s2(nsym, 0);
    for (size_t i = 0; i < nsym; i++) s1[i] = ((i / 5) & 1) ? 0x01 : 0xFF;          // dotting on the second carrier, third muted
    amps_fwd_params p;
    std::memset(&p, 0, sizeof p);
    p.samp_rate = 10e6; p.symrate = 100e3; p.max_deviation = 8000; p.device = 0; p.ncarriers = 3;
    p.carrier_freq[0] = 0; p.carrier_freq[1] = 60e3; p.carrier_freq[2] = 90e3;
    p.lpf_transition[0] = 5e3; p.lpf_transition[1] = 3e3; p.lpf_transition[2] = 3e3;
    p.out_scale = 0.5f; p.max_samples = (uint32_t)(nsym * 100);
    amps_fwd *h = NULL;
    if (amps_fwd_create(&p, &h) != AMPS_OK) { std::fprintf(stderr, "%s\n", amps_b200_last_error()); return 4; }
    std::vector<float> out(2 * nsym * 100);
    const uint8_t *syms[3] = {s0.data(), s1.data(), s2.data()};
    size_t half = nsym / 2;                                                         // two calls: exercises the carried history
    const bool voice = argc > 4 && !std::strcmp(argv[4], "voice");
    if (voice) {                                                                    // nsym must be a multiple of 50 here
        amps_fwd_voice_params vp = {1, 2, 16000.0, 8e3, 75e-6, 6000.0, 0.05};
        if (amps_fwd_enable_voice(h, &vp) != AMPS_OK) { std::fprintf(stderr, "%s\n", amps_b200_last_error()); return 7; }
        std::vector<float> audio(nsym * 4 / 25);
        for (size_t i = 0; i < audio.size(); i++) audio[i] = 0.3f * std::sin(0.17f * (float)i);
        half -= half % 25;
        if (amps_fwd_work_voice(h, syms, audio.data(), half, 0, out.data()) != AMPS_OK) return 5;
        const uint8_t *syms2[3] = {s0.data() + half, s1.data() + half, s2.data() + half};
        if (amps_fwd_work_voice(h, syms2, audio.data() + half * 4 / 25, nsym - half, 1, out.data() + 2 * half * 100) != AMPS_OK) {
            std::fprintf(stderr, "%s\n", amps_b200_last_error());
            return 6;
        }
    } else {
    if (amps_fwd_work(h, syms, half, out.data()) != AMPS_OK) return 5;
    const uint8_t *syms2[3] = {s0.data() + half, s1.data() + half, s2.data() + half};
    if (amps_fwd_work(h, syms2, nsym - half, out.data() + 2 * half * 100) != AMPS_OK) return 6;
    }
    amps_fwd_destroy(h);
    std::ofstream(argv[3], std::ios::binary).write(reinterpret_cast<const char *>(out.data()), (std::streamsize)(out.size() * sizeof(float)));
    std::printf("{\"samples\": %zu}\n", nsym * 100);
    return 0;
}

// forward_iq block: FOCC with an injected word pair, FVC alert train unmuted half-way, IQ to a file
static int run_txblock(int argc, char **argv) {
    if (argc < 4) return 2;
    const size_t nbits = std::strtoull(argv[2], NULL, 10);
    forward_iq::sptr tx = forward_iq::make(false, 0);
    probe pr;
    gr::msg_connect(*tx, "command_out", pr, "command_out");
    // page a mobile: word 1 + word 2 on both streams (what command_processor / recc_decode would send)
    unsigned char w1[28] = {0,1,0,0, 0,0,0,1,0,0,1,0,0,0,1,1,0,1,0,0,0,1,0,1,0,1,1,0}, w2[28] = {1,0,1,1, 0,1,0,1,0,1,0,1,0,1, 0, 0,0,0,0,0, 0,0,0, 0,0,0,0,0};
    tx->dispatch_msg("focc_words", pmt::make_tuple(pmt::from_long(3), pmt::from_long(2), pmt::mp(w1, 28), pmt::mp(w2, 28)));
    unsigned char alert[28] = {1,0,1,1,0,1,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1};
    tx->dispatch_msg("fvc_words", pmt::make_tuple(pmt::from_long(1), pmt::mp(alert, 28), pmt::from_uint64(2)));
    std::vector<std::complex<float>> out(nbits * 1000), buf(700 * 1000);
    gr_vector_const_void_star noin;
    gr_vector_void_star outs(1);
    size_t done = 0;
    while (done < nbits) {
        if (done >= nbits / 2) tx->dispatch_msg("fvc_mute", pmt::from_bool(false));
        size_t want = nbits - done < 700 ? nbits - done : 700;
        if (done < nbits / 2 && done + want > nbits / 2) want = nbits / 2 - done;
        outs[0] = buf.data();
        const int r = tx->work((int)(want * 1000), noin, outs);
        if (r <= 0 || r % 1000) return 3;
        std::memcpy(&out[done * 1000], buf.data(), (size_t)r * sizeof(std::complex<float>));
        done += (size_t)r / 1000;
    }
    std::ofstream(argv[3], std::ios::binary).write(reinterpret_cast<const char *>(out.data()), (std::streamsize)(out.size() * sizeof(std::complex<float>)));
    for (size_t i = 0; i < pr.lines.size(); i++) std::printf("%s\n", pr.lines[i].c_str());
    return 0;
}

// command_processor wired as grc/ampsbs.grc:4404-4412: every command is one PDU on "commands"
// The one real concurrency hazard of the block surface: another block's thread delivering focc_words while the scheduler
// thread is inside work() (lib/focc_impl.cc:567-580 guards its queue for that).  Thread B posts nmsgs two-word messages
// (stream BOTH; word i carries its own number), thread A keeps calling work(); the stream must stay a valid sequence of
// frames with every injected word in it exactly once, in order (checked by tests/test_host_gpu.py).
static int run_threads(int argc, char **argv) {
    if (argc < 5) return 2;
    const size_t total = std::strtoull(argv[2], NULL, 10);
    const int nmsgs = std::atoi(argv[3]);
    focc::sptr blk = focc::make(100000, false);
    std::atomic<bool> failed(false);
    std::atomic<size_t> produced(0);
    std::thread poster([&]() {
        for (int i = 0; i < nmsgs && !failed; i++) {
            unsigned char w[2][28];
            for (int k = 0; k < 2; k++) {
                const unsigned v = 0x5A00000u | (unsigned)(2 * i + k);               // 28 bits, MSB first
                for (int b = 0; b < 28; b++) w[k][b] = (unsigned char)((v >> (27 - b)) & 1u);
            }
            blk->dispatch_msg("focc_words", pmt::make_tuple(pmt::from_long(3), pmt::from_long(2), pmt::mp(w[0], 28), pmt::mp(w[1], 28)));
            // pace the posts over the first half of the run (a frame is 4630 bytes, a message two frames, and only the 15 filler
            // slots of a 19-frame superframe carry queued frames): the queue has drained when the run ends
            while (!failed && produced.load() < (size_t)(i + 1) * (total / (size_t)(2 * nmsgs + 4))) std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
    });
    std::vector<unsigned char> out, buf(1 << 16);
    gr_vector_const_void_star in;
    gr_vector_void_star outs(1);
    unsigned long long lcg = 99;
    while (out.size() < total) {
        lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
        int n = 1 + (int)((lcg >> 33) % 3000);
        if ((size_t)n > total - out.size()) n = (int)(total - out.size());
        outs[0] = buf.data();
        const int r = blk->work(n, in, outs);
        if (r < 0) { failed = true; break; }
        out.insert(out.end(), buf.begin(), buf.begin() + r);
        produced = out.size();
    }
    produced = total * 2;
    poster.join();
    if (failed) return 3;
    std::ofstream(argv[4], std::ios::binary).write(reinterpret_cast<const char *>(out.data()), (std::streamsize)out.size());
    std::printf("{\"bytes\": %zu, \"messages\": %d}\n", out.size(), nmsgs);
    return 0;
}

// malformed word messages: dropped with a warning (the reference asserts, lib/focc_impl.cc:523-532), never read past the tuple
static int run_badmsg() {
    focc::sptr f = focc::make(100000, false);
    fvc::sptr v = fvc::make(100000);
    unsigned char w[28] = {0};
    pmt::pmt_t good = pmt::mp(w, 28), shortb = pmt::mp(w, 5);
    f->dispatch_msg("focc_words", pmt::make_tuple(pmt::from_long(3), pmt::from_long(7), good));            // announces 7, carries 1
    f->dispatch_msg("focc_words", pmt::make_tuple(pmt::from_long(3), pmt::from_long(-2), good));           // negative count
    f->dispatch_msg("focc_words", pmt::make_tuple(pmt::from_long(3), pmt::from_long(1), shortb));          // short blob
    f->dispatch_msg("focc_words", pmt::make_tuple(pmt::from_long(3), pmt::from_long(1), good, good));      // more elements than words
    v->dispatch_msg("fvc_words", pmt::make_tuple(pmt::from_long(3), good));
    v->dispatch_msg("fvc_words", pmt::make_tuple(pmt::from_long(-1), good));
    v->dispatch_msg("fvc_words", pmt::make_tuple(pmt::from_long(1), shortb));
    // nothing was queued: the FOCC stream is the plain superframe, the FVC is still idle (buffer untouched)
    std::vector<unsigned char> a(4630), b(64, 0x77);
    gr_vector_const_void_star in;
    gr_vector_void_star outs(1);
    outs[0] = b.data();
    const int r = v->work(64, in, outs);
    bool untouched = true;
    for (size_t i = 0; i < b.size(); i++) untouched = untouched && b[i] == 0x77;
    std::printf("{\"fvc_work\": %d, \"fvc_buffer_untouched\": %s}\n", r, untouched ? "true" : "false");
    return 0;
}

// K carriers (-160 kHz + 30 kHz * k) demodulating ONE uploaded buffer through the batched C-ABI entry points; prints the bursts
// per channel as JSON lines (channel, demod_index, MIN) -- compared with stand-alone handles / the oracle by the tests, and a
// compact all-kernels scenario for compute-sanitizer.
struct batch_probe { std::vector<std::string> lines; };
static void on_batch_burst(int channel, const amps_burst *b, void *user) {
    char buf[160];
    std::snprintf(buf, sizeof buf, "{\"channel\": %d, \"demod_index\": %llu, \"min\": \"%.10s\", \"valid\": %d}", channel,
                  (unsigned long long)b->demod_index, b->decoded.min, (int)(b->decoded.valid[0] + b->decoded.valid[1] + b->decoded.valid[2] +
                  b->decoded.valid[3] + b->decoded.valid[4] + b->decoded.valid[5] + b->decoded.valid[6]));
    static_cast<batch_probe *>(user)->lines.push_back(buf);
}
static int run_batch(int argc, char **argv) {
    if (argc < 6) return 2;
    const size_t nsamples = std::strtoull(argv[3], NULL, 10);
    const int K = std::atoi(argv[4]);
    const size_t chunk = std::strtoull(argv[5], NULL, 10);
    std::vector<float> iq(2 * nsamples);
    std::ifstream f(argv[2], std::ios::binary);
    f.read(reinterpret_cast<char *>(iq.data()), (std::streamsize)(iq.size() * sizeof(float)));
    if (!f) { std::fprintf(stderr, "short read of %s\n", argv[2]); return 4; }
    std::vector<amps_recc_iq *> hs((size_t)K, NULL);
    for (int k = 0; k < K; k++) {
        amps_recc_iq_params p;
        std::memset(&p, 0, sizeof p);
        p.samp_rate = 10e6; p.center_freq = -160e3 + 30e3 * k; p.device = 0; p.max_samples = (uint32_t)chunk;
        if (amps_recc_iq_create(&p, &hs[(size_t)k]) != AMPS_OK) { std::fprintf(stderr, "%s\n", amps_b200_last_error()); return 5; }
    }
    amps_recc_iq_batch *b = NULL;
    if (amps_recc_iq_batch_create(hs.data(), K, 0, &b) != AMPS_OK) { std::fprintf(stderr, "%s\n", amps_b200_last_error()); return 6; }
    batch_probe pr;
    for (size_t pos = 0; pos < nsamples; pos += chunk) {
        const size_t n = std::min(chunk, nsamples - pos);
        if (amps_recc_iq_batch_work_shared(b, iq.data() + 2 * pos, n, &on_batch_burst, &pr) != AMPS_OK) { std::fprintf(stderr, "%s\n", amps_b200_last_error()); return 7; }
    }
    for (size_t i = 0; i < pr.lines.size(); i++) std::printf("%s\n", pr.lines[i].c_str());
    amps_recc_iq_batch_destroy(b);
    for (int k = 0; k < K; k++) amps_recc_iq_destroy(hs[(size_t)k]);
    return 0;
}

static int run_cmd(int argc, char **argv) {
    command_processor::sptr cp = command_processor::make();
    probe pr;
    const char *ports[] = {"focc_words", "fvc_words", "audio_mute", "fvc_mute", "debug_output"};
    for (size_t i = 0; i < sizeof(ports) / sizeof(ports[0]); i++) gr::msg_connect(*cp, ports[i], pr, ports[i]);
    for (int i = 2; i < argc; i++) {
        pr.lines.clear();
        cp->dispatch_msg("commands", pmt::cons(pmt::make_dict(), pmt::init_u8vector(std::strlen(argv[i]), (const uint8_t *)argv[i])));
        std::printf("[");
        for (size_t k = 0; k < pr.lines.size(); k++) std::printf("%s%s", k ? ", " : "", pr.lines[k].c_str());
        std::printf("]\n");
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: qa_blocks focc|loop|fwd ...\n"); return 2; }
    try {
        if (!std::strcmp(argv[1], "focc")) return run_focc(argc, argv);
        if (!std::strcmp(argv[1], "loop")) return run_loop(argc, argv);
        if (!std::strcmp(argv[1], "fwd")) return run_fwd(argc, argv);
        if (!std::strcmp(argv[1], "txblock")) return run_txblock(argc, argv);
        if (!std::strcmp(argv[1], "cmd")) return run_cmd(argc, argv);
        if (!std::strcmp(argv[1], "threads")) return run_threads(argc, argv);
        if (!std::strcmp(argv[1], "badmsg")) return run_badmsg();
        if (!std::strcmp(argv[1], "batch")) return run_batch(argc, argv);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "qa_blocks: %s\n", e.what());
        return 10;
    }
    return 2;
}
