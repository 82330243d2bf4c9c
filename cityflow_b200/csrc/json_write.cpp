// JSON output pieces (declared in replay.h): number and string printers.
//
// putJsonNumberLikeRapidjson restates the double printer of the reference's JSON writer -- rapidjson's internal dtoa, i.e.
// Grisu2 (F. Loitsch, "Printing floating-point numbers quickly and accurately with integers", PLDI 2010) followed by
// milo's Prettify layout -- so that an archive written here carries, digit for digit, the text the reference's
// Archive::dump (archive.cpp:153-177, utility.cpp:116-127) writes for the same doubles.  That matters beyond looks: the
// reference READS numbers with rapidjson's default (not full-precision) parser, which maps two different 17-digit
// spellings of one double to doubles one ulp apart, and Grisu2 does not always emit the shortest / closest spelling
// (for roughly one in ten of the 17-digit values an archive holds).  Same digits in => same doubles in the reference after loadFromFile.
//
// The 87 cached powers of ten Grisu2 multiplies with (10^k for k = -348, -340, ... 340 as 64-bit significand + binary
// exponent, rounded to nearest) are computed here at first use with exact integer arithmetic instead of being tabulated.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "replay.h"

namespace cfb {

namespace {

// ---- exact powers of ten -> 64-bit significand, binary exponent ----
struct Big {   // little-endian base 2^32, only what the table needs
    std::vector<uint32_t> w;
    int bits() const {
        for (size_t i = w.size(); i-- > 0;)
            if (w[i]) return (int) (i * 32 + 32 - __builtin_clz(w[i]));
        return 0;
    }
    void mulSmall(uint32_t m) {
        uint64_t carry = 0;
        for (auto &x : w) { const uint64_t p = (uint64_t) x * m + carry; x = (uint32_t) p; carry = p >> 32; }
        if (carry) w.push_back((uint32_t) carry);
    }
    bool bit(int i) const { return (size_t) (i >> 5) < w.size() && ((w[i >> 5] >> (i & 31)) & 1u); }
    void shl1(bool in) {
        uint32_t carry = in ? 1u : 0u;
        for (auto &x : w) { const uint32_t out = x >> 31; x = (x << 1) | carry; carry = out; }
        if (carry) w.push_back(carry);
    }
    bool geq(const Big &o) const {
        const size_t n = std::max(w.size(), o.w.size());
        for (size_t i = n; i-- > 0;) {
            const uint32_t a = i < w.size() ? w[i] : 0, b = i < o.w.size() ? o.w[i] : 0;
            if (a != b) return a > b;
        }
        return true;
    }
    void sub(const Big &o) {   // requires *this >= o
        int64_t borrow = 0;
        for (size_t i = 0; i < w.size(); ++i) {
            int64_t d = (int64_t) w[i] - (i < o.w.size() ? o.w[i] : 0) - borrow;
            borrow = d < 0;
            if (d < 0) d += (int64_t) 1 << 32;
            w[i] = (uint32_t) d;
        }
    }
};

struct DiyFp {
    uint64_t f = 0;
    int e = 0;
};

DiyFp powerOfTen(int k) {
    Big p;
    p.w.push_back(1);
    for (int i = 0; i < (k < 0 ? -k : k); ++i) p.mulSmall(10);
    unsigned __int128 top = 0;   // 65 leading bits of the value, the last one for rounding
    int e;
    if (k >= 0) {
        const int nb = p.bits();
        for (int i = 0; i < 65; ++i) top = (top << 1) | (unsigned) (nb - 1 - i >= 0 && p.bit(nb - 1 - i));
        e = nb - 64;
    } else {
        // 2^s / 10^|k| with s = bits(10^|k|) + 63 lies in (2^63, 2^64]: restoring division of 2^(s+1), bit by bit
        const int s = p.bits() + 63;
        Big r;
        r.w.push_back(0);
        for (int i = s + 1; i >= 0; --i) {
            r.shl1(i == s + 1);
            top <<= 1;
            if (r.geq(p)) { r.sub(p); top |= 1; }
        }
        e = -s;
    }
    unsigned __int128 f = (top + 1) >> 1;
    if (f >> 64) { f >>= 1; e += 1; }
    DiyFp out;
    out.f = (uint64_t) f;
    out.e = e;
    return out;
}

const DiyFp &cachedPower(int index) {   // 10^(-348 + 8 * index)
    static const std::vector<DiyFp> table = [] {
        std::vector<DiyFp> t;
        for (int i = 0; i < 87; ++i) t.push_back(powerOfTen(-348 + 8 * i));
        return t;
    }();
    return table[index];
}

// ---- Grisu2 ----
constexpr uint64_t kHidden = 1ull << 52;

DiyFp fromDouble(double d) {
    uint64_t u;
    memcpy(&u, &d, 8);
    const int biased = (int) ((u >> 52) & 0x7ff);
    const uint64_t frac = u & (kHidden - 1);
    DiyFp r;
    if (biased) { r.f = frac + kHidden; r.e = biased - 0x3ff - 52; }
    else { r.f = frac; r.e = 1 - 0x3ff - 52; }
    return r;
}
DiyFp mul(DiyFp a, DiyFp b) {
    const unsigned __int128 p = (unsigned __int128) a.f * b.f;
    uint64_t h = (uint64_t) (p >> 64);
    if ((uint64_t) p & (1ull << 63)) ++h;   // round the discarded half up
    DiyFp r;
    r.f = h;
    r.e = a.e + b.e + 64;
    return r;
}
DiyFp normalize(DiyFp a) {
    const int s = __builtin_clzll(a.f);
    a.f <<= s;
    a.e -= s;
    return a;
}
void boundaries(DiyFp v, DiyFp &minus, DiyFp &plus) {
    DiyFp pl;
    pl.f = (v.f << 1) + 1;
    pl.e = v.e - 1;
    while (!(pl.f & (kHidden << 1))) { pl.f <<= 1; pl.e--; }
    pl.f <<= 64 - 52 - 2;
    pl.e -= 64 - 52 - 2;
    DiyFp mi;
    if (v.f == kHidden) { mi.f = (v.f << 2) - 1; mi.e = v.e - 2; }
    else { mi.f = (v.f << 1) - 1; mi.e = v.e - 1; }
    mi.f <<= mi.e - pl.e;
    mi.e = pl.e;
    minus = mi;
    plus = pl;
}
void grisuRound(char *buffer, int len, uint64_t delta, uint64_t rest, uint64_t tenKappa, uint64_t wpW) {
    while (rest < wpW && delta - rest >= tenKappa && (rest + tenKappa < wpW || wpW - rest > rest + tenKappa - wpW)) {
        buffer[len - 1]--;
        rest += tenKappa;
    }
}
int decimalDigits32(uint32_t n) {
    if (n < 10) return 1;
    if (n < 100) return 2;
    if (n < 1000) return 3;
    if (n < 10000) return 4;
    if (n < 100000) return 5;
    if (n < 1000000) return 6;
    if (n < 10000000) return 7;
    if (n < 100000000) return 8;
    return 9;   // the integral part has at most 9 digits here (Loitsch's choice of the cached power)
}
constexpr uint32_t kPow10[] = {1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000, 1000000000};

void digitGen(DiyFp W, DiyFp Mp, uint64_t delta, char *buffer, int *len, int *K) {
    DiyFp one;
    one.f = 1ull << -Mp.e;
    one.e = Mp.e;
    const uint64_t wpW = Mp.f - W.f;
    uint32_t p1 = (uint32_t) (Mp.f >> -one.e);
    uint64_t p2 = Mp.f & (one.f - 1);
    int kappa = decimalDigits32(p1);
    *len = 0;
    while (kappa > 0) {
        const uint32_t div = kPow10[kappa - 1];
        const uint32_t d = p1 / div;
        p1 %= div;
        if (d || *len) buffer[(*len)++] = (char) ('0' + d);
        kappa--;
        const uint64_t tmp = ((uint64_t) p1 << -one.e) + p2;
        if (tmp <= delta) {
            *K += kappa;
            grisuRound(buffer, *len, delta, tmp, (uint64_t) kPow10[kappa] << -one.e, wpW);
            return;
        }
    }
    for (;;) {
        p2 *= 10;
        delta *= 10;
        const char d = (char) (p2 >> -one.e);
        if (d || *len) buffer[(*len)++] = (char) ('0' + d);
        p2 &= one.f - 1;
        kappa--;
        if (p2 < delta) {
            *K += kappa;
            const int index = -kappa;
            grisuRound(buffer, *len, delta, p2, one.f, wpW * (index < 9 ? kPow10[index] : 0));
            return;
        }
    }
}
void grisu2(double value, char *buffer, int *length, int *K) {
    const DiyFp v = fromDouble(value);
    DiyFp wm, wp;
    boundaries(v, wm, wp);
    // smallest cached power that brings the product's exponent into Grisu's window
    const double dk = (-61 - wp.e) * 0.30102999566398114 + 347;   // (alpha - e - 1) * log10(2) + |k_min| - 1
    int k = (int) dk;
    if (dk - k > 0.0) k++;
    const int index = (k >> 3) + 1;
    *K = -(-348 + index * 8);
    const DiyFp c = cachedPower(index);
    const DiyFp W = mul(normalize(v), c);
    DiyFp Wp = mul(wp, c), Wm = mul(wm, c);
    Wm.f++;
    Wp.f--;
    digitGen(W, Wp, Wp.f - Wm.f, buffer, length, K);
}

// "2.0", "12.34", "0.001234", "1.234e33", "1e-7": digits + decimal exponent -> text
void layout(std::string &s, const char *digits, int length, int k) {
    const int kk = length + k;   // 10^(kk-1) <= v < 10^kk
    if (0 <= k && kk <= 21) {
        s.append(digits, length);
        s.append((size_t) k, '0');
        s += ".0";
    } else if (0 < kk && kk <= 21) {
        s.append(digits, kk);
        s.push_back('.');
        s.append(digits + kk, length - kk);
    } else if (-6 < kk && kk <= 0) {
        s += "0.";
        s.append((size_t) -kk, '0');
        s.append(digits, length);
    } else {
        s.push_back(digits[0]);
        if (length > 1) { s.push_back('.'); s.append(digits + 1, length - 1); }
        s.push_back('e');
        s += std::to_string(kk - 1);
    }
}

}  // namespace

void putJsonNumberLikeRapidjson(std::string &s, double v) {
    if (v == 0) { s += std::signbit(v) ? "-0.0" : "0.0"; return; }
    if (v < 0) { s.push_back('-'); v = -v; }
    char digits[32];
    int length = 0, K = 0;
    grisu2(v, digits, &length, &K);
    layout(s, digits, length, K);
}

// Shortest digit string that parses back to v (std::to_chars), in the same layout.
void putJsonNumber(std::string &s, double v) {
    if (v == 0) { s += "0.0"; return; }
    if (v < 0) { s.push_back('-'); v = -v; }
    char buf[40];
    const auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
    char digits[24];
    int length = 0, expo = 0;
    const char *q = buf;
    for (; q < r.ptr && *q != 'e'; ++q)
        if (*q != '.') digits[length++] = *q;
    if (q < r.ptr) {
        ++q;
        const bool neg = *q == '-';
        if (*q == '-' || *q == '+') ++q;
        for (; q < r.ptr; ++q) expo = expo * 10 + (*q - '0');
        if (neg) expo = -expo;
    }
    layout(s, digits, length, expo + 1 - length);
}

void putJsonString(std::string &s, const std::string &v) {
    s.push_back('"');
    for (unsigned char c : v) {
        if (c == '"' || c == '\\') { s.push_back('\\'); s.push_back((char) c); }
        else if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04X", c); s.append(b); }
        else s.push_back((char) c);
    }
    s.push_back('"');
}

}  // namespace cfb
