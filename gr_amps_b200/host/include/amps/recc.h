// recc.h -- same public surface as the reference's include/amps/recc.h:28
#pragma once
#include <amps/api.h>
#include <gnuradio/sync_block.h>
namespace gr { namespace amps {
class AMPS_API recc : virtual public gr::sync_block {
public:
    typedef std::shared_ptr<recc> sptr;
    static sptr make();
};
}}
