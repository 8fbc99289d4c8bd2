// command_processor.h -- same public surface as the reference's include/amps/command_processor.h:15
#pragma once
#include <amps/api.h>
#include <gnuradio/block.h>
namespace gr { namespace amps {
class AMPS_API command_processor : virtual public gr::block {
public:
    typedef std::shared_ptr<command_processor> sptr;
    static sptr make();
};
}}
