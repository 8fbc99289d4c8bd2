// stand-in for <boost/algorithm/string/trim.hpp>: in-place trim of isspace characters (lib/command_processor_impl.cc:112)
#pragma once
#include <cctype>
#include <string>
namespace boost {
inline void trim(std::string &s) {
    size_t b = 0, e = s.size();
    while (b < e && std::isspace((unsigned char)s[b])) ++b;
    while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
    s = s.substr(b, e - b);
}
}  // namespace boost
