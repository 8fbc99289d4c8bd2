// forwarder: the Boost stand-in lives with the GNU Radio stand-in (gr_amps_b200/host/gr_shim/boost_standin.h).  TEST INFRASTRUCTURE.
#pragma once
#include "../../gr_amps_b200/host/gr_shim/boost_standin.h"
