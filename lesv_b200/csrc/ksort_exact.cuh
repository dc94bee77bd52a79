// ksort_exact.cuh -- the permutation-exact restatement of klib's introsort as the reference uses it.
//
// lesv sorts records with comparators that look at part of the record only (hits by score, hits by
// subject id, windows by score ...), so WHICH of two equal-key records comes
// first is decided by the sorting algorithm itself and is visible in the output.  Every such site
// calls ks_introsort_* (reference src/corelib/ksort.h:180-229): median-of-three quicksort that leaves
// runs of <= 16 unsorted, an explicit stack, a comb sort when the depth budget runs out (:158-179) and
// one final insertion sort over the whole array (:150-157).  This file follows that control flow step
// by step with array indices instead of pointers; it runs on the host and inside kernels.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define KSX_HD __host__ __device__
#else
#define KSX_HD
#endif

namespace lesvsort {

template <class T, class Less>
KSX_HD void insertion_sort(T* a, long lo, long hi_excl, Less lt)          // ksort.h:150-157
{
    for (long i = lo + 1; i < hi_excl; ++i)
        for (long j = i; j > lo && lt(a[j], a[j - 1]); --j) {
            const T x = a[j]; a[j] = a[j - 1]; a[j - 1] = x;
        }
}

template <class T, class Less>
KSX_HD void comb_sort(T* a, size_t n, Less lt)                            // ksort.h:158-179
{
    const double shrink = 1.2473309501039786540366528676643;
    size_t gap = n;
    bool swapped;
    do {
        if (gap > 2) {
            gap = (size_t)(gap / shrink);
            if (gap == 9 || gap == 10) gap = 11;
        }
        swapped = false;
        for (size_t i = 0; i + gap < n; ++i) {
            const size_t j = i + gap;
            if (lt(a[j], a[i])) {
                const T x = a[i]; a[i] = a[j]; a[j] = x;
                swapped = true;
            }
        }
    } while (swapped || gap > 2);
    if (gap != 1) insertion_sort(a, 0, (long)n, lt);
}

template <class T, class Less>
KSX_HD void introsort(T* a, size_t n, Less lt)                            // ksort.h:180-229
{
    if (n < 1) return;
    if (n == 2) {
        if (lt(a[1], a[0])) { const T x = a[0]; a[0] = a[1]; a[1] = x; }
        return;
    }
    int d = 2;
    while ((1ul << d) < n) ++d;
    struct Frame { long left, right; int depth; };
    Frame stack[64];      // the larger side is pushed and the smaller one continued: depth <= log2(n)
    int top = 0;
    long s = 0, t = (long)n - 1;
    d <<= 1;
    for (;;) {
        if (s < t) {
            if (--d == 0) {
                comb_sort(a + s, (size_t)(t - s + 1), lt);
                t = s;
                continue;
            }
            long i = s, j = t, k = i + ((j - i) >> 1) + 1;
            if (lt(a[k], a[i])) {
                if (lt(a[k], a[j])) k = j;
            } else k = lt(a[j], a[i]) ? i : j;
            const T rp = a[k];
            if (k != t) { const T x = a[k]; a[k] = a[t]; a[t] = x; }
            for (;;) {
                do ++i; while (lt(a[i], rp));
                do --j; while (i <= j && lt(rp, a[j]));
                if (j <= i) break;
                const T x = a[i]; a[i] = a[j]; a[j] = x;
            }
            { const T x = a[i]; a[i] = a[t]; a[t] = x; }
            if (i - s > t - i) {
                if (i - s > 16) { stack[top].left = s; stack[top].right = i - 1; stack[top].depth = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { stack[top].left = i + 1; stack[top].right = t; stack[top].depth = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == 0) {
                insertion_sort(a, 0, (long)n, lt);
                return;
            }
            --top;
            s = stack[top].left; t = stack[top].right; d = stack[top].depth;
        }
    }
}

}  // namespace lesvsort
