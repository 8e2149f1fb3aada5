// Stand-in for the Boost subset the reference uses. Written for this repository; see the README.md of oracle/refshim.
// boost::regex / regex_replace with a "$1" format (rt.hpp:2297-2298) -> std::regex (ECMAScript, same format syntax)
#pragma once
#include <regex>
#include <string>
namespace boost {
typedef std::regex regex;
inline std::string regex_replace(const std::string& s, const regex& re, const std::string& fmt) { return std::regex_replace(s, re, fmt); }
}
