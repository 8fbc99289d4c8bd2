// api.h -- export macro, as include/amps/api.h:4-10 of the reference
#pragma once
#include <gnuradio/attributes.h>
#ifdef gnuradio_amps_EXPORTS
#  define AMPS_API __GR_ATTR_EXPORT
#else
#  define AMPS_API __GR_ATTR_IMPORT
#endif
