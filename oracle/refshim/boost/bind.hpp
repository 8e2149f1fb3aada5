// Stand-in for the Boost subset the reference uses. Written for this repository; see the README.md of oracle/refshim.
// boost::bind(&io_service::run, &io) (basic.hpp:110) -> std::bind
#pragma once
#include <functional>
namespace boost { using std::bind; }
