// Stand-in for the Boost subset the reference uses. Written for this repository; see the README.md of oracle/refshim.
// (included by include/nanogi/basic.hpp:48-52 but nothing of it is used)
#pragma once
