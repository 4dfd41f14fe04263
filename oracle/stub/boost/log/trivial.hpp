/* oracle/stub/boost/log/trivial.hpp — TEST INFRASTRUCTURE.
 * Boost is not installed in this image; the reference's hot-path headers only use
 * BOOST_LOG_TRIVIAL for one informational line (simulation_parameters.cuh:241).  This stub
 * swallows the stream and pulls in the std headers real Boost.Log brings transitively. */
#pragma once
#include <cstring>
#include <iostream>
#include <iterator>
#include <string>
struct swo_null_log {
    template <class T> swo_null_log &operator<<(const T &) { return *this; }
    swo_null_log &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }
};
#define BOOST_LOG_TRIVIAL(lvl) swo_null_log()
