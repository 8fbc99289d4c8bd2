// fvc.h -- same public surface as the reference's include/amps/fvc.h:34
#pragma once
#include <amps/api.h>
#include <gnuradio/sync_block.h>
namespace gr { namespace amps {
class AMPS_API fvc : virtual public gr::sync_block {
public:
    typedef std::shared_ptr<fvc> sptr;
    static sptr make(unsigned long symrate);
};
}}
