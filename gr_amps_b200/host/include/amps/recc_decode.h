// recc_decode.h -- same public surface as the reference's include/amps/recc_decode.h:28
#pragma once
#include <amps/api.h>
#include <gnuradio/block.h>
namespace gr { namespace amps {
class AMPS_API recc_decode : virtual public gr::block {
public:
    typedef boost::shared_ptr<recc_decode> sptr;     // GNU Radio 3.7's block pointer type (include/amps/focc.h:24 of the reference)
    static sptr make();
};
}}
