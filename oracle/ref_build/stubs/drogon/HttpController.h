// Compile-only stand-in for Drogon: the controller class declaration and drogon::app() have to parse; nothing runs.
#pragma once
#include <cstdlib>
#include <functional>
#include <memory>
namespace drogon {
struct HttpRequest;
struct HttpResponse;
using HttpRequestPtr = std::shared_ptr<HttpRequest>;
using HttpResponsePtr = std::shared_ptr<HttpResponse>;
template <class T> class HttpController {};
struct App {
    App &addListener(const char *, int) { std::abort(); }
    void run() { std::abort(); }
};
inline App &app() { static App a; return a; }
} // namespace drogon
#define METHOD_LIST_BEGIN
#define METHOD_LIST_END
#define ADD_METHOD_TO(...)
