// Compile-only stand-in for nlohmann::json: just enough surface for the reference's HTTP helpers to compile.
// None of these is on the path of the functions oracle/ref_build/ref_shim_*.cpp calls; calling them aborts.
#pragma once
#include <cstdlib>
#include <string>
namespace nlohmann {
class json {
  public:
    json() = default;
    static json parse(const std::string &) { std::abort(); }
    struct proxy {
        template <class T> proxy &operator=(const T &) { std::abort(); }
    };
    proxy operator[](const char *) { std::abort(); }
    json at(const char *) const { std::abort(); }
    std::string dump() const { std::abort(); }
    template <class T> T get() const { std::abort(); }
};
} // namespace nlohmann
