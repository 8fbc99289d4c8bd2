// stand-in for <gnuradio/attributes.h> (include/amps/api.h:4-10 of the reference)
#pragma once
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
