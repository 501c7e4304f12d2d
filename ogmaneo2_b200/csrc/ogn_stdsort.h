// ogn_stdsort.h — the permutation libstdc++'s std::sort (GCC 13, bits/stl_algo.h / stl_heap.h: introsort with a
// median-of-three pivot, insertion sort below 16 elements, heap sort when the depth limit 2*floor(log2 n) runs out) produces
// for std::pair<float, int> elements under the comparator `lhs.first > rhs.first`, callable on host and device.
//
// Why this exists: the reference's ImageEncoder ranks the cells of a hidden column with exactly that call
// (ImageEncoder.cpp:15-17, :56). std::sort is not stable, so for EQUAL activations the rank order is whatever this
// algorithm leaves — and at realistic sizes ties do occur (about one tied pair per step among 4096 columns of 32 cells),
// which changes which cell loses more of its resource. Replaying the algorithm keeps the resources and weights
// bit-identical to the reference. The restatement is pinned by tests/test_stdsort.py, which compares it with the real
// std::sort on millions of tie-heavy arrays (tests/csrc/stdsort_check.cpp).
//
// Elements are (key[i], idx[i]) pairs kept in two parallel arrays; "moves" move both.
#pragma once

#if defined(__CUDACC__)
#define OGN_SORT_HD __host__ __device__ inline
#else
#define OGN_SORT_HD static inline
#endif

namespace ogn_stdsort {

struct Elem { float k; int i; };
OGN_SORT_HD Elem get(const float *key, const int *idx, int p) { Elem e; e.k = key[p]; e.i = idx[p]; return e; }
OGN_SORT_HD void put(float *key, int *idx, int p, const Elem &e) { key[p] = e.k; idx[p] = e.i; }
OGN_SORT_HD void swp(float *key, int *idx, int a, int b) { const Elem t = get(key, idx, a); put(key, idx, a, get(key, idx, b)); put(key, idx, b, t); }
// comp(a, b): a goes before b ("largest in front")
OGN_SORT_HD bool cmp(float a, float b) { return a > b; }

// std::__unguarded_linear_insert / __insertion_sort / __unguarded_insertion_sort
OGN_SORT_HD void unguarded_linear_insert(float *key, int *idx, int last) {
    const Elem val = get(key, idx, last);
    int next = last - 1;
    while (cmp(val.k, key[next])) { put(key, idx, last, get(key, idx, next)); last = next; --next; }
    put(key, idx, last, val);
}
OGN_SORT_HD void insertion_sort(float *key, int *idx, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (cmp(key[i], key[first])) {
            const Elem val = get(key, idx, i);
            for (int j = i; j > first; --j) put(key, idx, j, get(key, idx, j - 1)); // move_backward(first, i, i + 1)
            put(key, idx, first, val);
        } else
            unguarded_linear_insert(key, idx, i);
    }
}
OGN_SORT_HD void unguarded_insertion_sort(float *key, int *idx, int first, int last) {
    for (int i = first; i != last; ++i) unguarded_linear_insert(key, idx, i);
}

// stl_heap.h: __push_heap / __adjust_heap / __make_heap / __pop_heap / __sort_heap on [first, first + len)
OGN_SORT_HD void push_heap(float *key, int *idx, int first, int hole, int top, const Elem &value) {
    int parent = (hole - 1) / 2;
    while (hole > top && cmp(key[first + parent], value.k)) {
        put(key, idx, first + hole, get(key, idx, first + parent));
        hole = parent;
        parent = (hole - 1) / 2;
    }
    put(key, idx, first + hole, value);
}
OGN_SORT_HD void adjust_heap(float *key, int *idx, int first, int hole, int len, const Elem &value) {
    const int top = hole;
    int second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (cmp(key[first + second], key[first + second - 1])) second--;
        put(key, idx, first + hole, get(key, idx, first + second));
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        put(key, idx, first + hole, get(key, idx, first + second - 1));
        hole = second - 1;
    }
    push_heap(key, idx, first, hole, top, value);
}
OGN_SORT_HD void heap_sort(float *key, int *idx, int first, int last) { // __partial_sort(first, last, last)
    const int len = last - first;
    if (len >= 2) { // __make_heap
        int parent = (len - 2) / 2;
        while (true) {
            const Elem value = get(key, idx, first + parent);
            adjust_heap(key, idx, first, parent, len, value);
            if (parent == 0) break;
            parent--;
        }
    }
    int l = last;
    while (l - first > 1) { // __sort_heap: __pop_heap(first, l - 1, l - 1)
        --l;
        const Elem value = get(key, idx, l);
        put(key, idx, l, get(key, idx, first));
        adjust_heap(key, idx, first, 0, l - first, value);
    }
}

OGN_SORT_HD int floor_log2(int n) { int k = 0; while (n > 1) { n >>= 1; ++k; } return k; }

// std::sort(first, last, comp) for n elements; idx[] normally starts as 0..n-1
OGN_SORT_HD void sort_desc(float *key, int *idx, int n) {
    if (n <= 0) return;
    // __introsort_loop with its tail recursion unrolled into an explicit stack of (first, last, depth) ranges: the right
    // part of every partition is sorted first (the recursive call), then the loop continues on the left part
    int stF[64], stL[64], stD[64], sp = 0;
    stF[0] = 0; stL[0] = n; stD[0] = 2 * floor_log2(n); sp = 1;
    while (sp > 0) {
        --sp;
        int first = stF[sp], last = stL[sp], depth = stD[sp];
        while (last - first > 16) {
            if (depth == 0) {
#ifdef OGN_STDSORT_COUNT_HEAP
                ++OGN_STDSORT_COUNT_HEAP; // test hook: how often the depth limit was reached
#endif
                heap_sort(key, idx, first, last);
                break;
            }
            --depth;
            // __unguarded_partition_pivot: median of (first + 1, mid, last - 1) moved to first
            const int mid = first + (last - first) / 2, a = first + 1, b = mid, c = last - 1;
            if (cmp(key[a], key[b])) {
                if (cmp(key[b], key[c])) swp(key, idx, first, b);
                else if (cmp(key[a], key[c])) swp(key, idx, first, c);
                else swp(key, idx, first, a);
            } else if (cmp(key[a], key[c])) swp(key, idx, first, a);
            else if (cmp(key[b], key[c])) swp(key, idx, first, c);
            else swp(key, idx, first, b);
            int lo = first + 1, hi = last; // __unguarded_partition(first + 1, last, pivot = first)
            while (true) {
                while (cmp(key[lo], key[first])) ++lo;
                --hi;
                while (cmp(key[first], key[hi])) --hi;
                if (!(lo < hi)) break;
                swp(key, idx, lo, hi);
                ++lo;
            }
            const int cut = lo;
            // recursive call on [cut, last) happens BEFORE the loop continues on [first, cut): since the two ranges are
            // disjoint and nothing else reads them in between, the order of processing does not change the result; push
            // the right part and keep looping on the left
            stF[sp] = cut; stL[sp] = last; stD[sp] = depth; ++sp;
            last = cut;
        }
    }
    // __final_insertion_sort
    if (n > 16) { insertion_sort(key, idx, 0, 16); unguarded_insertion_sort(key, idx, 16, n); }
    else insertion_sort(key, idx, 0, n);
}

} // namespace ogn_stdsort
