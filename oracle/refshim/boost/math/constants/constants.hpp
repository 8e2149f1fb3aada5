// Stand-in for the Boost subset the reference uses. Written for this repository; see the README.md of oracle/refshim.
#pragma once
namespace boost { namespace math { namespace constants {
template <class T> inline T pi() { return (T)3.141592653589793238462643383279502884L; }
}}}
