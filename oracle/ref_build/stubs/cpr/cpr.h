// Compile-only stand-in for cpr (HTTP client): never called by the shim.
#pragma once
#include <cstdlib>
#include <string>
namespace cpr {
struct Url { explicit Url(const std::string &) {} };
struct Body { explicit Body(const std::string &) {} };
struct Response { std::string text; long status_code = 0; };
inline Response Get(const Url &) { std::abort(); }
inline Response Post(const Url &, const Body &) { std::abort(); }
} // namespace cpr
