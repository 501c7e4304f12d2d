// tests/csrc/stdsort_check.cpp — compares ogn_stdsort::sort_desc (ogmaneo2_b200/csrc/ogn_stdsort.h) with the real
// std::sort on std::pair<float, int> under the comparator the reference's ImageEncoder uses (ImageEncoder.cpp:15-17), on
// tie-heavy random arrays of every length up to 64 and on adversarial patterns. Prints "ok <cases>" or the first mismatch.
#include <algorithm>
#include <cstdio>
#include <random>
#include <utility>
#include <vector>

static long g_heapSorts = 0;
#define OGN_STDSORT_COUNT_HEAP g_heapSorts
#include "../../ogmaneo2_b200/csrc/ogn_stdsort.h"

static bool pairfiCompare(const std::pair<float, int> &lhs, const std::pair<float, int> &rhs) { return lhs.first > rhs.first; }

static bool check(const std::vector<float> &v) {
    const int n = (int)v.size();
    std::vector<std::pair<float, int>> a((size_t)n);
    std::vector<float> key(v);
    std::vector<int> idx((size_t)n);
    for (int i = 0; i < n; i++) { a[(size_t)i] = std::make_pair(v[(size_t)i], i); idx[(size_t)i] = i; }
    std::sort(a.begin(), a.end(), pairfiCompare);
    ogn_stdsort::sort_desc(key.data(), idx.data(), n);
    for (int i = 0; i < n; i++)
        if (a[(size_t)i].second != idx[(size_t)i] || a[(size_t)i].first != key[(size_t)i]) {
            std::printf("mismatch n=%d at %d:", n, i);
            for (int j = 0; j < n; j++) std::printf(" %g", v[(size_t)j]);
            std::printf("\n");
            return false;
        }
    return true;
}

int main(int argc, char **argv) {
    const long cases = argc > 1 ? std::atol(argv[1]) : 2000000;
    std::mt19937 rng(12345);
    long done = 0;
    for (long c = 0; c < cases; c++) {
        const int n = 1 + (int)(rng() % 64);
        const int distinct = 1 + (int)(rng() % (c % 3 == 0 ? 3 : (c % 3 == 1 ? 8 : 1000)));
        std::vector<float> v((size_t)n);
        for (auto &x : v) x = -(float)(rng() % (unsigned)distinct) * 0.37f;
        if (c % 11 == 0) std::sort(v.begin(), v.end());
        if (c % 13 == 0) std::sort(v.begin(), v.end(), [](float x, float y) { return x > y; });
        if (c % 17 == 0) for (int i = 0; i < n; i++) v[(size_t)i] = (float)((i * 7919) % (n / 2 + 1)); // organ-pipe-ish
        if (!check(v)) return 1;
        done++;
    }
    // adversarial inputs (McIlroy, "A killer adversary for quicksort", 1999): a comparator that decides the values while the
    // real std::sort runs, so that every pivot lands near an end; the frozen values then drive both sorts to the depth limit
    // and into the heap-sort fallback. Both signs, every length from 17 to 64.
    for (int n = 17; n <= 64; n++)
        for (int sign = -1; sign <= 1; sign += 2) {
            std::vector<int> val((size_t)n, n - 1), ptr((size_t)n);
            const int gas = n - 1;
            int nsolid = 0, candidate = 0;
            for (int i = 0; i < n; i++) ptr[(size_t)i] = i;
            std::sort(ptr.begin(), ptr.end(), [&](int x, int y) {
                if (val[(size_t)x] == gas && val[(size_t)y] == gas) {
                    if (x == candidate) val[(size_t)x] = nsolid++;
                    else val[(size_t)y] = nsolid++;
                }
                if (val[(size_t)x] == gas) candidate = x;
                else if (val[(size_t)y] == gas) candidate = y;
                return sign * val[(size_t)x] > sign * val[(size_t)y];
            });
            std::vector<float> v((size_t)n);
            for (int i = 0; i < n; i++) v[(size_t)i] = (float)(sign * val[(size_t)i]);
            if (!check(v)) return 1;
            for (int i = 0; i < n; i++) v[(size_t)i] = (float)(sign * (val[(size_t)i] / 2)); // the same with ties
            if (!check(v)) return 1;
            done += 2;
        }
    std::printf("ok %ld heap_sorts %ld\n", done, g_heapSorts);
    return 0;
}
