// boost_standin.h -- TEST INFRASTRUCTURE.  The few Boost names the gr-amps sources reach through the GNU Radio
// headers (boost::shared_ptr, boost::bind(&T::handler, this, _1), boost::mutex::scoped_lock), mapped onto the C++
// standard library so that the UNMODIFIED reference sources under /root/reference/lib compile here (oracle/Makefile,
// target _ref).  Boost is not installed in this image.
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
