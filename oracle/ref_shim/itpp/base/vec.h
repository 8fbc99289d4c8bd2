// stand-in for <itpp/base/vec.h>: bvec lives in the BCH stand-in
#pragma once
#include <itpp/comm/bch.h>
