// forward_iq.h -- NEW composite source block: the whole forward path of grc/ampsbs.grc in one block.  It replaces
//   amps.focc -> char_to_float -> frequency_modulator_fc -> pfb.interpolator_ccf            (FOCC leg, carrier 0)
//   amps.fvc  -> char_to_float -> frequency_modulator_fc -> pfb.interpolator_ccf -> mute -> x e^{j 2 pi 60k t}   (FVC leg)
//   add -> multiply_const 0.5                                        (grc/ampsbs.grc:4422,4548,4476,4650,4530,4560)
// with the fused B200 path at 10 MS/s (a third, muted carrier slot stands for the audio-only +90 kHz leg).  Message
// ports keep the reference names: in "focc_words" (lib/focc_impl.cc:127), "fvc_words" (lib/fvc_impl.cc:62),
// "fvc_mute" (mute_xx set_mute, grc/ampsbs.grc:1555-1601); out "command_out" ("fvc off", lib/fvc_impl.cc:163-171).
#pragma once
#include <amps/api.h>
#include <gnuradio/sync_block.h>
#include <complex>
namespace gr { namespace amps {
class AMPS_API forward_iq : virtual public gr::sync_block {
public:
    typedef boost::shared_ptr<forward_iq> sptr;     // GNU Radio 3.7's block pointer type (include/amps/focc.h:24 of the reference)
    static sptr make(bool aggressive_registration, int device = 0);
};
}}
