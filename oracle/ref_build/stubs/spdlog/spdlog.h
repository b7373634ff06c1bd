// Compile-only stand-in for spdlog (absent here): the reference logs through SPDLOG_INFO / SPDLOG_ERROR with
// fmt-style arguments.  The arguments are handed to a hook so that benchmark_results' numbers (which the reference
// only logs) can be read back; nothing else of spdlog is used by the reference's sources.  Test infrastructure.
#pragma once
// (the real header pulls in most of the standard library; the reference relies on that for <cmath>, <array>, ...)
#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <stdexcept>
#include <cstring>
#include <filesystem>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>
void pf_ref_log_line(const std::string &line);
template <class... A>
inline void pf_ref_log(const char *fmt, const A &...args) {
    std::ostringstream os;
    os.precision(9);
    os << fmt;
    ((os << " | " << args), ...);
    pf_ref_log_line(os.str());
}
#define SPDLOG_INFO(...) pf_ref_log(__VA_ARGS__)
#define SPDLOG_ERROR(...) pf_ref_log(__VA_ARGS__)
