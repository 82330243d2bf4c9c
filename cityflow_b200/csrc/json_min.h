// Minimal JSON DOM for the config / roadnet / flow files.
//
// The reference parses these files with rapidjson (extern/rapidjson, default parse flags) and
// every derived length / cross distance feeds the FP64 hot path, so number conversion must give
// the *same double, bit for bit*.  rapidjson's default ("normal precision") path is NOT a
// correctly rounded strtod: it accumulates up to 2^53 of the digits in an integer, the rest in a
// double, and scales once by a table power of ten (reader.h ParseNumber, internal/strtod.h
// StrtodNormalPrecision/FastPath).  number() below restates that published algorithm.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace cfb {

struct JsonError : std::runtime_error {
    explicit JsonError(const std::string &m) : std::runtime_error(m) {}
};

class Json {
public:
    enum Kind { Null, False, True, Int, Uint, Int64, Uint64, Double, String, Array, Object };
    Kind kind = Null;
    int64_t i = 0;       // Int / Int64
    uint64_t u = 0;      // Uint / Uint64
    double d = 0.0;      // Double
    std::string s;       // String
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;  // insertion order; first match wins (FindMember)

    bool isObject() const { return kind == Object; }
    bool isArray() const { return kind == Array; }
    bool isString() const { return kind == String; }
    bool isBool() const { return kind == True || kind == False; }
    bool isNumber() const { return kind >= Int && kind <= Double; }
    // rapidjson IsInt(): value fits a signed 32-bit int
    bool isInt() const {
        if (kind == Int) return true;
        if (kind == Uint) return u <= 0x7fffffffu;
        return false;
    }
    int asInt() const { return kind == Int ? (int) i : (int) u; }
    unsigned asUint() const { return kind == Int ? (unsigned) i : (unsigned) u; }
    bool asBool() const { return kind == True; }
    double asDouble() const {
        switch (kind) {
            case Double: return d;
            case Int: case Int64: return (double) i;
            case Uint: case Uint64: return (double) u;
            default: return 0.0;
        }
    }
    const Json *find(const std::string &key) const {
        if (kind != Object) return nullptr;
        for (const auto &kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }

    static Json parseFile(const std::string &path, bool *opened = nullptr);
    static Json parse(const char *p, size_t n);
};

namespace detail {

inline double pow10tab(int n) {
    static double tab[309];
    static bool init = false;
    if (!init) {
        char buf[16];
        for (int k = 0; k <= 308; ++k) {  // correctly rounded literals, as a compiler would emit
            snprintf(buf, sizeof buf, "1e%d", k);
            tab[k] = strtod(buf, nullptr);
        }
        init = true;
    }
    return tab[n];
}

inline double scale10(double sig, int exp) {
    if (exp < -308) return 0.0;
    if (exp >= 0) return sig * pow10tab(exp);
    return sig / pow10tab(-exp);
}

struct Parser {
    const char *p, *end;
    int line = 1;
    [[noreturn]] void fail(const char *what) {
        throw JsonError(std::string("Json parsing error (") + what + ") at line " + std::to_string(line));
    }
    int peek() const { return p < end ? (unsigned char) *p : -1; }
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) {
            if (*p == '\n') ++line;
            ++p;
        }
    }
    bool eat(char c) {
        if (p < end && *p == c) { ++p; return true; }
        return false;
    }
    static bool dig(int c) { return c >= '0' && c <= '9'; }

    Json number() {
        Json out;
        bool minus = eat('-');
        unsigned i = 0;
        uint64_t i64 = 0;
        bool use64 = false, useDouble = false;
        int sigDigits = 0;
        double d = 0.0;
        if (peek() == '0') {
            ++p;
        } else if (peek() >= '1' && peek() <= '9') {
            i = (unsigned) (*p++ - '0');
            const unsigned lim = minus ? 214748364u : 429496729u;
            const char last = minus ? '8' : '5';
            while (dig(peek())) {
                if (i >= lim && (i != lim || *p > last)) { i64 = i; use64 = true; break; }
                i = i * 10 + (unsigned) (*p++ - '0');
                ++sigDigits;
            }
        } else {
            fail("invalid value");
        }
        if (use64) {
            const uint64_t lim = minus ? 0x0CCCCCCCCCCCCCCCull : 0x1999999999999999ull;
            const char last = minus ? '8' : '5';
            while (dig(peek())) {
                if (i64 >= lim && (i64 != lim || *p > last)) { d = (double) i64; useDouble = true; break; }
                i64 = i64 * 10 + (unsigned) (*p++ - '0');
                ++sigDigits;
            }
        }
        if (useDouble)
            while (dig(peek())) d = d * 10 + (*p++ - '0');
        int expFrac = 0;
        if (eat('.')) {
            if (!dig(peek())) fail("missing fraction");
            if (!useDouble) {
                if (!use64) i64 = i;
                while (dig(peek())) {
                    if (i64 > 0x1FFFFFFFFFFFFFull) break;  // 2^53 - 1
                    i64 = i64 * 10 + (unsigned) (*p++ - '0');
                    --expFrac;
                    if (i64 != 0) ++sigDigits;
                }
                d = (double) i64;
                useDouble = true;
            }
            while (dig(peek())) {
                if (sigDigits < 17) {
                    d = d * 10.0 + (*p++ - '0');
                    --expFrac;
                    if (d > 0.0) ++sigDigits;
                } else {
                    ++p;
                }
            }
        }
        int exp = 0;
        if (eat('e') || eat('E')) {
            if (!useDouble) { d = (double) (use64 ? i64 : i); useDouble = true; }
            bool expMinus = false;
            if (eat('+')) {} else if (eat('-')) expMinus = true;
            if (!dig(peek())) fail("missing exponent");
            exp = *p++ - '0';
            if (expMinus) {
                int maxExp = (expFrac + 2147483639) / 10;
                while (dig(peek())) {
                    exp = exp * 10 + (*p++ - '0');
                    if (exp > maxExp)
                        while (dig(peek())) ++p;
                }
            } else {
                int maxExp = 308 - expFrac;
                while (dig(peek())) {
                    exp = exp * 10 + (*p++ - '0');
                    if (exp > maxExp) fail("number too big");
                }
            }
            if (expMinus) exp = -exp;
        }
        if (useDouble) {
            int pw = exp + expFrac;
            if (pw < -308) { d = scale10(d, -308); d = scale10(d, pw + 308); }
            else d = scale10(d, pw);
            if (d > 1.7976931348623157e308) fail("number too big");
            out.kind = Json::Double;
            out.d = minus ? -d : d;
        } else if (use64) {
            if (minus) { out.kind = Json::Int64; out.i = (int64_t) (~i64 + 1); }
            else { out.kind = Json::Uint64; out.u = i64; }
        } else {
            if (minus) { out.kind = Json::Int; out.i = (int32_t) (~i + 1); }
            else { out.kind = Json::Uint; out.u = i; }
        }
        return out;
    }

    static void utf8(std::string &s, unsigned cp) {
        if (cp < 0x80) s += (char) cp;
        else if (cp < 0x800) { s += (char) (0xC0 | (cp >> 6)); s += (char) (0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) {
            s += (char) (0xE0 | (cp >> 12)); s += (char) (0x80 | ((cp >> 6) & 0x3F)); s += (char) (0x80 | (cp & 0x3F));
        } else {
            s += (char) (0xF0 | (cp >> 18)); s += (char) (0x80 | ((cp >> 12) & 0x3F));
            s += (char) (0x80 | ((cp >> 6) & 0x3F)); s += (char) (0x80 | (cp & 0x3F));
        }
    }
    unsigned hex4() {
        unsigned v = 0;
        for (int k = 0; k < 4; ++k) {
            int c = peek();
            ++p;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string str() {
        std::string s;
        if (!eat('"')) fail("expected string");
        for (;;) {
            if (p >= end) fail("unterminated string");
            char c = *p++;
            if (c == '"') break;
            if (c == '\\') {
                if (p >= end) fail("bad escape");
                char e = *p++;
                switch (e) {
                    case '"': s += '"'; break;
                    case '\\': s += '\\'; break;
                    case '/': s += '/'; break;
                    case 'b': s += '\b'; break;
                    case 'f': s += '\f'; break;
                    case 'n': s += '\n'; break;
                    case 'r': s += '\r'; break;
                    case 't': s += '\t'; break;
                    case 'u': {
                        unsigned cp = hex4();
                        if (cp >= 0xD800 && cp <= 0xDBFF && p + 1 < end && p[0] == '\\' && p[1] == 'u') {
                            p += 2;
                            unsigned lo = hex4();
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        }
                        utf8(s, cp);
                        break;
                    }
                    default: fail("bad escape");
                }
            } else {
                s += c;
            }
        }
        return s;
    }
    Json value() {
        ws();
        Json v;
        int c = peek();
        if (c == '{') {
            ++p;
            v.kind = Json::Object;
            ws();
            if (eat('}')) return v;
            for (;;) {
                ws();
                std::string k = str();
                ws();
                if (!eat(':')) fail("missing colon");
                v.obj.emplace_back(std::move(k), value());
                ws();
                if (eat(',')) continue;
                if (eat('}')) break;
                fail("missing comma or }");
            }
        } else if (c == '[') {
            ++p;
            v.kind = Json::Array;
            ws();
            if (eat(']')) return v;
            for (;;) {
                v.arr.push_back(value());
                ws();
                if (eat(',')) continue;
                if (eat(']')) break;
                fail("missing comma or ]");
            }
        } else if (c == '"') {
            v.kind = Json::String;
            v.s = str();
        } else if (c == 't') {
            if (end - p < 4 || memcmp(p, "true", 4)) fail("invalid value");
            p += 4; v.kind = Json::True;
        } else if (c == 'f') {
            if (end - p < 5 || memcmp(p, "false", 5)) fail("invalid value");
            p += 5; v.kind = Json::False;
        } else if (c == 'n') {
            if (end - p < 4 || memcmp(p, "null", 4)) fail("invalid value");
            p += 4; v.kind = Json::Null;
        } else {
            v = number();
        }
        return v;
    }
};

}  // namespace detail

inline Json Json::parse(const char *p, size_t n) {
    detail::Parser ps{p, p + n};
    ps.ws();
    if (ps.p == ps.end) ps.fail("document empty");
    Json v = ps.value();
    ps.ws();
    if (ps.p != ps.end) ps.fail("trailing characters");
    return v;
}

inline Json Json::parseFile(const std::string &path, bool *opened) {
    FILE *fp = fopen(path.c_str(), "rb");
    if (opened) *opened = fp != nullptr;
    if (!fp) return Json();
    std::string buf;
    char tmp[1 << 16];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, fp)) > 0) buf.append(tmp, n);
    fclose(fp);
    return parse(buf.data(), buf.size());
}

}  // namespace cfb
