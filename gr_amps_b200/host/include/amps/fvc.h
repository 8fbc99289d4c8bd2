// fvc.h -- same public surface as the reference's include/amps/fvc.h:34
#pragma once
#include <amps/api.h>
#include <gnuradio/sync_block.h>
namespace gr { namespace amps {
class AMPS_API fvc : virtual public gr::sync_block {
public:
    typedef boost::shared_ptr<fvc> sptr;     // GNU Radio 3.7's block pointer type (include/amps/focc.h:24 of the reference)
    static sptr make(unsigned long symrate);
};
}}
