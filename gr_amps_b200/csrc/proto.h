// proto.h -- AMPS protocol helpers of the host side of libamps_b200 (plain C++): BCH(40,28) parity,
// the 28-bit FOCC/FVC word layouts and the 463-slot FOCC frame.  Behaviour follows
// lib/focc_impl.cc:156-218,252-381, lib/amps_packet.cc:26-95, lib/utils.cc:101-108 of the reference.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace amps {

using Word28 = std::array<uint8_t, 28>;
constexpr uint8_t kSlotBI = 2;         // busy/idle slot marker in a frame's slot table

void expandbits(uint8_t *out, int nbits, uint64_t val);
std::array<uint8_t, 40> bch_encode_40_28(const uint8_t *info28);

Word28 overhead_word_1(unsigned dcc, unsigned sid, bool ep, bool auth, bool pci, unsigned nawc);
Word28 overhead_word_2(unsigned dcc, bool s, bool e, bool regh, bool regr, unsigned dtx, unsigned nminusone, bool rcf,
                       bool cpa, unsigned cmax, bool end);
Word28 control_filler_word();
Word28 access_type_parameters_global_action(unsigned dcc, bool end);
Word28 registration_increment_global_action(unsigned dcc, unsigned regincr, bool end);
Word28 registration_id(unsigned dcc, unsigned long regid, bool end);
Word28 focc_word1(bool multiword, unsigned dcc, uint64_t min1);
Word28 focc_word2_general(uint64_t min2, unsigned msg_type, unsigned ordq, unsigned order);
Word28 fvc_word1_general(unsigned pscc, unsigned msg_type, unsigned ordq, unsigned order);
Word28 focc_word2_voice_channel(unsigned scc, uint64_t min2, unsigned vmac, unsigned chan);

// MIN digits <-> MIN1 (24 bit) / MIN2 (10 bit): lib/amps_packet.h:305-349.  parse_min needs exactly 10 digits
// (the reference accepts 1..10 but reads min[0..9] regardless; the short cases are out-of-bounds reads there).
uint64_t compute_min_3(char d1, char d2, char d3);
bool parse_min(const std::string &min, uint64_t &min1, uint64_t &min2);

// 463 slots: [BI] dotting(10) [BI] sync(11), then 5 x (A, B) words as 4 x ([BI] + 10 bits)
std::array<uint8_t, 463> focc_frame_slots(const uint8_t *word_a28, const uint8_t *word_b28);
// 1032 bits: 101 dotting + 11 x (sync 11 + word 40) with 37-bit dotting between repeats
std::vector<uint8_t> fvc_word_train(const uint8_t *word28);

}  // namespace amps
