// pf_json.hpp — the JSON envelope of the reference's endpoints, as far as they need it: objects whose values are
// numbers, strings or nested arrays of numbers (ref: src/server/controllers/Query.cc:34-59 and
// src/client/client_lib.cpp:104-119 use nlohmann::json for exactly these shapes), plus base64 for the ciphertext
// bodies of the additive encrypted endpoint.  nlohmann-json is not in this image; where it is available it can
// replace this codec without touching the handler bodies (pf_query_handlers.hpp) or the client (pf_client.hpp).
// Shared by the server-side handlers and the client-side request / response helpers; no CUDA, no engine.
#pragma once
#include <charconv>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <string_view>
#include <type_traits>
#include <vector>

namespace prefhetch::handlers {

// ---- JSON, as far as these endpoints need it ------------------------------------------------------------
namespace json {

inline void skip_ws(std::string_view s, size_t &i) {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) i++;
}

// end (one past) of the value starting at i: string, array, object, or bare token
inline size_t value_end(std::string_view s, size_t i) {
    if (i >= s.size()) throw std::runtime_error("json: unexpected end of input");
    if (s[i] == '"') {
        for (size_t j = i + 1; j < s.size(); j++) {
            if (s[j] == '\\') j++;
            else if (s[j] == '"') return j + 1;
        }
        throw std::runtime_error("json: unterminated string");
    }
    if (s[i] == '[' || s[i] == '{') {
        int depth = 0;
        for (size_t j = i; j < s.size(); j++) {
            if (s[j] == '"') j = value_end(s, j) - 1;
            else if (s[j] == '[' || s[j] == '{') depth++;
            else if (s[j] == ']' || s[j] == '}') {
                if (--depth == 0) return j + 1;
            }
        }
        throw std::runtime_error("json: unbalanced brackets");
    }
    size_t j = i;
    while (j < s.size() && s[j] != ',' && s[j] != ']' && s[j] != '}' && s[j] != ' ' && s[j] != '\n' && s[j] != '\t' && s[j] != '\r') j++;
    if (j == i) throw std::runtime_error("json: value expected");
    return j;
}

// top-level object -> key -> raw text of the value (views into `body`)
inline std::map<std::string, std::string_view> object(std::string_view body) {
    std::map<std::string, std::string_view> out;
    size_t i = 0;
    skip_ws(body, i);
    if (i >= body.size() || body[i] != '{') throw std::runtime_error("json: object expected");
    i++;
    skip_ws(body, i);
    if (i < body.size() && body[i] == '}') return out;
    for (;;) {
        skip_ws(body, i);
        if (i >= body.size() || body[i] != '"') throw std::runtime_error("json: key expected");
        const size_t ke = value_end(body, i);
        std::string key(body.substr(i + 1, ke - i - 2));
        i = ke;
        skip_ws(body, i);
        if (i >= body.size() || body[i] != ':') throw std::runtime_error("json: ':' expected");
        i++;
        skip_ws(body, i);
        const size_t ve = value_end(body, i);
        out.emplace(std::move(key), body.substr(i, ve - i));
        i = ve;
        skip_ws(body, i);
        if (i < body.size() && body[i] == ',') {
            i++;
            continue;
        }
        if (i < body.size() && body[i] == '}') return out;
        throw std::runtime_error("json: ',' or '}' expected");
    }
}

// the views point INTO the body: a temporary string must not be parsed (object(handler(...)) would leave them dangling)
template <class S, std::enable_if_t<std::is_same_v<std::remove_cv_t<S>, std::string>, int> = 0>
std::map<std::string, std::string_view> object(S &&) = delete;

inline std::string_view at(const std::map<std::string, std::string_view> &o, const char *key) {
    auto it = o.find(key);
    if (it == o.end()) throw std::runtime_error(std::string("json: key '") + key + "' not found"); // nlohmann: out_of_range
    return it->second;
}

template <class T> T number(std::string_view tok) {
    T v{};
    const char *b = tok.data(), *e = tok.data() + tok.size();
    if (b < e && *b == '+') b++;
    auto [p, ec] = std::from_chars(b, e, v);
    if (ec != std::errc() || p != e) throw std::runtime_error("json: bad number '" + std::string(tok) + "'");
    return v;
}

// flat array of numbers: [a, b, ...]
template <class T> std::vector<T> vector(std::string_view s) {
    std::vector<T> out;
    size_t i = 0;
    skip_ws(s, i);
    if (i >= s.size() || s[i] != '[') throw std::runtime_error("json: array expected");
    i++;
    skip_ws(s, i);
    if (i < s.size() && s[i] == ']') return out;
    for (;;) {
        skip_ws(s, i);
        const size_t ve = value_end(s, i);
        out.push_back(number<T>(s.substr(i, ve - i)));
        i = ve;
        skip_ws(s, i);
        if (i < s.size() && s[i] == ',') {
            i++;
            continue;
        }
        if (i < s.size() && s[i] == ']') return out;
        throw std::runtime_error("json: ',' or ']' expected");
    }
}

// array of equally long arrays of numbers -> row-major flat vector; rows / cols returned
template <class T> std::vector<T> matrix(std::string_view s, size_t &rows, size_t &cols) {
    std::vector<T> out;
    rows = cols = 0;
    size_t i = 0;
    skip_ws(s, i);
    if (i >= s.size() || s[i] != '[') throw std::runtime_error("json: array of arrays expected");
    i++;
    skip_ws(s, i);
    if (i < s.size() && s[i] == ']') return out;
    for (;;) {
        skip_ws(s, i);
        const size_t ve = value_end(s, i);
        const std::vector<T> row = vector<T>(s.substr(i, ve - i));
        if (rows == 0) cols = row.size();
        else if (row.size() != cols) throw std::runtime_error("json: rows of different lengths"); // std::array would not convert
        out.insert(out.end(), row.begin(), row.end());
        rows++;
        i = ve;
        skip_ws(s, i);
        if (i < s.size() && s[i] == ',') {
            i++;
            continue;
        }
        if (i < s.size() && s[i] == ']') return out;
        throw std::runtime_error("json: ',' or ']' expected");
    }
}

inline std::string string(std::string_view s) {
    size_t i = 0;
    skip_ws(s, i);
    if (i >= s.size() || s[i] != '"') throw std::runtime_error("json: string expected");
    const size_t e = value_end(s, i);
    std::string out;
    for (size_t j = i + 1; j + 1 < e; j++) {
        if (s[j] == '\\' && j + 2 < e) {
            const char c = s[++j];
            out.push_back(c == 'n' ? '\n' : c == 't' ? '\t' : c == 'r' ? '\r' : c); // \" \\ \/ and the common escapes
        } else {
            out.push_back(s[j]);
        }
    }
    return out;
}

template <class T> void put_number(std::string &out, T v) {
    char buf[40];
    auto [p, ec] = std::to_chars(buf, buf + sizeof(buf), v); // floats: shortest text that reads back to the same value
    (void)ec;
    out.append(buf, p);
}
template <class T> void put_vector(std::string &out, const T *v, size_t n) {
    out.push_back('[');
    for (size_t i = 0; i < n; i++) {
        if (i) out.push_back(',');
        put_number(out, v[i]);
    }
    out.push_back(']');
}
template <class T> void put_matrix(std::string &out, const T *v, size_t rows, size_t cols) {
    out.push_back('[');
    for (size_t r = 0; r < rows; r++) {
        if (r) out.push_back(',');
        put_vector(out, v + r * cols, cols);
    }
    out.push_back(']');
}

inline std::string base64_encode(const uint8_t *p, size_t n) {
    static const char *A = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    std::string out;
    out.reserve((n + 2) / 3 * 4);
    size_t i = 0;
    for (; i + 2 < n; i += 3) {
        const uint32_t v = (uint32_t)p[i] << 16 | (uint32_t)p[i + 1] << 8 | p[i + 2];
        out.push_back(A[v >> 18]);
        out.push_back(A[(v >> 12) & 63]);
        out.push_back(A[(v >> 6) & 63]);
        out.push_back(A[v & 63]);
    }
    if (i < n) {
        const uint32_t v = (uint32_t)p[i] << 16 | (i + 1 < n ? (uint32_t)p[i + 1] << 8 : 0);
        out.push_back(A[v >> 18]);
        out.push_back(A[(v >> 12) & 63]);
        out.push_back(i + 1 < n ? A[(v >> 6) & 63] : '=');
        out.push_back('=');
    }
    return out;
}

inline std::vector<uint8_t> base64_decode(std::string_view s) {
    static const auto table = [] {
        std::vector<int8_t> t(256, -1);
        const char *A = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; i++) t[(uint8_t)A[i]] = (int8_t)i;
        return t;
    }();
    std::vector<uint8_t> out;
    out.reserve(s.size() / 4 * 3);
    uint32_t acc = 0;
    int bits = 0;
    for (char c : s) {
        if (c == '=' || c == '\n' || c == '\r') continue;
        const int8_t v = table[(uint8_t)c];
        if (v < 0) throw std::runtime_error("base64: invalid character");
        acc = acc << 6 | (uint32_t)v;
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back((uint8_t)(acc >> bits));
        }
    }
    return out;
}

} // namespace json

} // namespace prefhetch::handlers
