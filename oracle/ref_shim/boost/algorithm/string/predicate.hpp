// stand-in for <boost/algorithm/string/predicate.hpp>: starts_with / istarts_with (lib/command_processor_impl.cc:97-110)
#pragma once
#include <cctype>
#include <string>
namespace boost {
inline bool starts_with(const std::string &s, const std::string &p) { return s.size() >= p.size() && s.compare(0, p.size(), p) == 0; }
inline bool istarts_with(const std::string &s, const std::string &p) {
    if (s.size() < p.size()) return false;
    for (size_t i = 0; i < p.size(); ++i)
        if (std::tolower((unsigned char)s[i]) != std::tolower((unsigned char)p[i])) return false;
    return true;
}
}  // namespace boost
