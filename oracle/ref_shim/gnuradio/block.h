// stand-in: the GNU Radio 3.7 block base classes (gr_amps_b200/host/gr_shim) + the Boost names GNU Radio's own
// headers would have pulled in.  TEST INFRASTRUCTURE (oracle/_ref build only).
#pragma once
#include "../boost_standin.h"
#include <cassert>
#include <cstdio>
#include <cstring>
#include <sys/types.h>
#include_next <gnuradio/sync_block.h>
