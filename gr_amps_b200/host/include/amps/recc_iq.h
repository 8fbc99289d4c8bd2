// recc_iq.h -- NEW sibling of amps.recc: complex IQ in (10 MS/s), "bursts" message port out.  It replaces
// freq_xlating_fir_filter_ccc -> quadrature_demod_cf -> clock_recovery_mm_ff -> binary_slicer_fb -> amps.recc
// (grc/ampsbs.grc:4656,4620,4506,4614,4602) with the fused B200 path; the blob it publishes has the same
// layout as the one recc_impl::work publishes (lib/recc_impl.cc:126).
#pragma once
#include <amps/api.h>
#include <gnuradio/sync_block.h>
#include <complex>
namespace gr { namespace amps {
class AMPS_API recc_iq : virtual public gr::sync_block {
public:
    typedef boost::shared_ptr<recc_iq> sptr;     // GNU Radio 3.7's block pointer type (include/amps/focc.h:24 of the reference)
    // mm_timing: symbol timing by the reference graph's clock_recovery_mm_ff -> binary_slicer_fb -> amps.recc tail
    // (AMPS_RX_TIMING_MM) instead of the feed-forward trigger detector
    // sc16: the input stream is interleaved int16 I,Q (item size 4: the USRP's wire format, uhd stream_args cpu_format
    // "sc16") instead of gr_complex; converted on the GPU as (float)int16 / 32768 (AMPS_RX_INPUT_SC16)
    static sptr make(double samp_rate, double center_freq, int device = 0, bool mm_timing = false, bool sc16 = false);
};
}}
