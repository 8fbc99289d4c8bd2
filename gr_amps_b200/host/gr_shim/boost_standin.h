// boost_standin.h -- part of the GNU Radio 3.7 STAND-IN (gr_shim), used only where real GNU Radio / Boost headers are absent.
// GNU Radio 3.7's public types are Boost types: block sptrs are boost::shared_ptr (include/amps/focc.h:24 of the reference,
// gnuradio::get_initial_sptr), handlers are bound with boost::bind(&T::handler, this, _1), locks are
// boost::mutex::scoped_lock.  Boost is not installed in this image, so those names are mapped onto the C++ standard library
// here; with real GNU Radio 3.7 on the include path this file is never seen and the same sources get the real Boost types.
#pragma once
#include <memory>
#include <mutex>

namespace boost {
using std::shared_ptr;

struct placeholder1 {};
// boost::bind(&C::method, this, _1) as used at lib/focc_impl.cc:128-130 and friends
template <class R, class C, class A, class T>
auto bind(R (C::*m)(A), T *self, placeholder1) {
    return [m, self](A a) { return (self->*m)(a); };
}

class mutex {
public:
    class scoped_lock {
    public:
        explicit scoped_lock(mutex &m) : d_g(m.d_m) {}
    private:
        std::lock_guard<std::mutex> d_g;
    };
private:
    std::mutex d_m;
};
}  // namespace boost

static const boost::placeholder1 _1 = boost::placeholder1();
