// command_processor.h -- operator-console block of the AMPS base station, host only (no GPU work).
//
// Drop-in for the reference block of the same name: class name, namespace, base class and make() signature are the
// ABI that the flowgraph (grc/amps_command_processor.xml: "amps.command_processor()") and SWIG bind to, so they are
// kept; everything behind them is in host/lib/blocks_impl.cc.
//
// Message ports (names are part of the wiring of grc/ampsbs.grc:4404-4470):
//   in   "commands"      PDU (dict . u8vector) carrying one text command
//   out  "focc_words"    tuple (stream = 3, 2, word1[28], word2[28])      -- "page <10-digit MIN>"
//   out  "fvc_words"     tuple (1, word[28])                               -- "fvc alert"
//   out  "fvc_mute" / "audio_mute"   bool                                  -- "fvc on" / "fvc off"
//   out  "debug_output"  PDU with a human-readable answer
// Commands and answers follow lib/command_processor_impl.cc:52-117; a MIN shorter than ten digits is refused
// ("invalid MIN entered") where the reference reads past the end of the string.
#pragma once
#include <amps/api.h>
#include <gnuradio/block.h>


namespace gr { namespace amps {

class AMPS_API command_processor : virtual public gr::block {
public:
    typedef boost::shared_ptr<command_processor> sptr;     // GNU Radio 3.7's block pointer type (include/amps/focc.h:24 of the reference)
    static sptr make();
};

}}  // namespace gr::amps
