// json_number_probe.cpp -- TEST INFRASTRUCTURE: csrc/json_write.cpp's restatement of rapidjson's double printer (Grisu2 +
// Prettify) against (a) rapidjson itself, when built with -DWITH_RAPIDJSON -I<reference>/extern/rapidjson/include
// (modes `live N` and `gen N out`), and (b) a committed list of (bits, text) pairs that rapidjson produced
// (tests/golden/rapidjson_dtoa.txt, mode `golden file`), which travels to machines without the reference.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#ifdef WITH_RAPIDJSON
#include "rapidjson/internal/dtoa.h"
#endif
#include "replay.h"

static std::vector<double> samples(long n, uint64_t seed) {
    std::mt19937_64 rng(seed);
    std::vector<double> v = {0.0, -0.0, 1.0, 1e21, 1e22, 1e-6, 1e-7, 5e-324, 1.7976931348623157e308, 2.2250738585072014e-308, 0.1,
                             123456789012345678.0, 16.67, 8.3333, 2147483647.0};
    std::uniform_real_distribution<double> U(0, 1);
    for (long i = 0; i < n; ++i) {
        uint64_t u = rng();
        double d;
        memcpy(&d, &u, 8);
        if (d == d && d - d == 0) v.push_back(d);                 // any finite bit pattern
        const double x = U(rng);                                  // the ranges an archive holds
        v.push_back(x * 300); v.push_back(x * 16.67); v.push_back(x * 1e-9); v.push_back((double) (rng() % 100000) + x);
        v.push_back(-x * 5); v.push_back((double) (rng() % 1000)); v.push_back((rng() % 1000) * 0.5); v.push_back(1.0 / (1 + rng() % 1000));
    }
    return v;
}

int main(int argc, char **argv) {
    if (argc >= 3 && !strcmp(argv[1], "golden")) {
        FILE *f = fopen(argv[2], "r");
        if (!f) { printf("cannot open %s\n", argv[2]); return 2; }
        char text[80];
        unsigned long long bits;
        long n = 0, bad = 0;
        while (fscanf(f, "%llx %79s", &bits, text) == 2) {
            double d;
            memcpy(&d, &bits, 8);
            std::string mine;
            cfb::putJsonNumberLikeRapidjson(mine, d);
            if (mine != text && bad++ < 10) printf("DIFF %016llx: golden %s ours %s\n", bits, text, mine.c_str());
            ++n;
        }
        fclose(f);
        printf("%s %ld values, %ld mismatches\n", bad ? "FAIL" : "OK", n, bad);
        return bad != 0;
    }
#ifdef WITH_RAPIDJSON
    if (argc >= 3 && (!strcmp(argv[1], "live") || !strcmp(argv[1], "gen"))) {
        const bool gen = !strcmp(argv[1], "gen");
        FILE *out = gen && argc >= 4 ? fopen(argv[3], "w") : nullptr;
        long bad = 0, n = 0, notShortest = 0;
        for (double d : samples(atol(argv[2]), gen ? 777 : 12345)) {
            char buf[64];
            char *e = rapidjson::internal::dtoa(d, buf);
            const std::string ref(buf, e);
            std::string mine, shortest;
            cfb::putJsonNumberLikeRapidjson(mine, d);
            cfb::putJsonNumber(shortest, d);
            if (mine != ref && bad++ < 10) printf("DIFF %a: rapidjson %s ours %s\n", d, ref.c_str(), mine.c_str());
            notShortest += shortest != ref && d != 0;
            if (out) { unsigned long long bits; memcpy(&bits, &d, 8); fprintf(out, "%016llx %s\n", bits, ref.c_str()); }
            ++n;
        }
        if (out) fclose(out);
        printf("%s %ld values, %ld mismatches, %ld where rapidjson's text is not the shortest one\n", bad ? "FAIL" : "OK", n, bad, notShortest);
        return bad != 0;
    }
#endif
    printf("usage: json_number_probe golden <file> | live <n> | gen <n> <out>   (live / gen need -DWITH_RAPIDJSON)\n");
    return 64;
}
