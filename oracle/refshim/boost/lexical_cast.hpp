// Stand-in for the Boost subset the reference uses. Written for this repository; see the README.md of oracle/refshim.
#pragma once
#include <sstream>
#include <string>
namespace boost {
template <class T, class S> inline T lexical_cast(const S& s) { std::stringstream ss; ss << s; T out; ss >> out; return out; }
}
