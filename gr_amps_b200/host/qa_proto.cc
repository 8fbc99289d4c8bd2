// qa_proto.cc -- CPU-only QA of the PRODUCT's host-side helpers (csrc/proto.cc, csrc/design.cc): prints JSON that
// tests/test_host_cpu.py compares with the golden KATs and with the oracle.  No CUDA involved.
#include "../csrc/design.h"
#include "../csrc/proto.h"

#include <cstdio>
#include <string>
#include <vector>

using namespace amps;

template <typename W> static std::string bits(const W &w, size_t n) {
    std::string s;
    for (size_t i = 0; i < n; i++) s += w[i] ? '1' : '0';
    return s;
}
static void word(const char *name, const Word28 &w, bool last = false) {
    const auto e = bch_encode_40_28(w.data());
    std::printf("  \"%s\": [\"%s\", \"%s\"]%s\n", name, bits(w, 28).c_str(), bits(e, 40).substr(28).c_str(), last ? "" : ",");
}
static void taps(const char *name, const std::vector<float> &t, bool last = false) {
    std::printf("  \"%s\": [", name);
    for (size_t i = 0; i < t.size(); i++) std::printf("%s%.17g", i ? ", " : "", (double)t[i]);
    std::printf("]%s\n", last ? "" : ",");
}

int main() {
    std::printf("{\n \"words\": {\n");
    word("OW1 nawc=3", overhead_word_1(0, 16, true, false, false, 3));
    word("OW1 nawc=4", overhead_word_1(0, 16, true, false, false, 4));
    word("OW2", overhead_word_2(0, true, true, true, true, 0, 23, true, true, 23, false));
    word("control filler", control_filler_word());
    word("access-type GA END=0", access_type_parameters_global_action(0, false));
    word("REGINCR=100 END=0", registration_increment_global_action(0, 100, false));
    word("REGID=0 END=1", registration_id(0, 0, true));
    word("REGID=500 END=1", registration_id(0, 500, true));
    word("FVC alert order scc=1", fvc_word1_general(1, 0, 0, 1));
    word("focc_word1", focc_word1(true, 0, 0xABCDE));
    word("focc_word2_general", focc_word2_general(0x155, 0, 0, 7));
    word("focc_word2_voice_channel", focc_word2_voice_channel(1, 0x2AA, 0, 355), true);
    std::printf(" },\n");
    const Word28 ow1 = overhead_word_1(0, 16, true, false, false, 3);
    const auto frame = focc_frame_slots(ow1.data(), ow1.data());
    std::string fs;
    for (size_t i = 0; i < frame.size(); i++) fs += (char)('0' + frame[i]);
    std::printf(" \"frame_slots\": \"%s\",\n", fs.c_str());
    const Word28 alert = fvc_word1_general(1, 0, 0, 1);
    const auto train = fvc_word_train(alert.data());
    std::printf(" \"fvc_train\": \"%s\",\n", bits(train, train.size()).c_str());
    std::printf(" \"fcw\": [%u, %u],\n", nco_fcw(-160e3, 10e6), nco_fcw(-160e3, 400e3));
    std::vector<float> cic;
    cic3_taps(25, cic);
    std::printf(" \"taps\": {\n");
    taps("gr_qa_firdes_low_pass", firdes_low_pass(1.0, 1.0, 0.4, 0.2, WIN_HAMMING));     // GNU Radio qa_firdes.py test_low_pass
    taps("lpf", firdes_low_pass(3.0, 400e3, 10e3, 4500.0, WIN_BLACKMAN));
    taps("focc_interp", firdes_low_pass(1.0, 400e3, 10e3, 5e3, WIN_HAMMING));
    taps("fvc_interp", firdes_low_pass(1.0, 400e3, 10e3, 3e3, WIN_HAMMING));
    taps("cic25", cic);
    taps("mmse", mmse_interp_table());
    {
        int per = 0;
        taps("voice_arb25", arb25_taps(firdes_low_pass(3.0, 400e3, 15e3, 6e3, WIN_BLACKMAN), per));
        double b[2], a[2];
        fm_preemph_taps(16000.0, 75e-6, -1.0, b, a);
        std::printf("  \"preemph\": [%.17g, %.17g, %.17g, %.17g],\n", b[0], b[1], a[0], a[1]);
        const std::vector<double> g = fm_preemph_impulse(16000.0, 75e-6, -1.0, 192);
        std::printf("  \"preemph_impulse\": [");
        for (size_t i = 0; i < g.size(); i++) std::printf("%s%.17g", i ? ", " : "", g[i]);
        std::printf("],\n");
    }
    taps("voice_lpf", firdes_low_pass(3.0, 400e3, 15e3, 6e3, WIN_BLACKMAN), true);
    std::printf(" }\n}\n");
    return 0;
}
