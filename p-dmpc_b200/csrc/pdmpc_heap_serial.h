// pdmpc_heap_serial.h — the reference's priority queue as ONE thread executes it.
//
// priority_queue_interface_mex.cpp:19-31 keeps std::priority_queue<(id, f)> with
// the comparator f_a > f_b; pop order among equal f is defined by libstdc++'s
// heap mechanics (stl_heap.h __push_heap :135-147, __adjust_heap :224-249,
// __pop_heap :254-262).  The two routines below leave the array in exactly the
// state those routines leave it in, so every later tie resolves identically.
//
// pop() does not walk the hole to a leaf and back like __adjust_heap +
// __push_heap do.  Along the descent path c_1, c_2, ... (c_t = the child
// libstdc++ would pick: the right one unless f_right > f_left) the heap order
// makes f non-decreasing, so the entries __push_heap moves back down are a
// suffix of the path: the net effect of the two library routines is "entries
// c_1..c_T move up one level, v lands on c_T" with T = the number of leading
// path entries with f <= v.f.  The walk below stops there.
// tests/test_heap_serial.py checks it against the reference's own MEX source
// compiled here (oracle/_ref/libpq_ref.so) on tie-dense sequences.
//
// Compiles as host code (tests) and device code (the lane-per-search kernel).
#pragma once

#if defined(__CUDACC__)
#define PDMPC_HD __host__ __device__ __forceinline__
#else
#define PDMPC_HD inline
#endif

namespace pdmpc {

// E: struct with a double member `f`.  H: random-access array of E (heap position i at H[i]).
template <class E, class H>
PDMPC_HD void heap_push_serial(H h, int len, const E &v) {
    int hole = len;
    while (hole > 0) {
        const int parent = (hole - 1) >> 1;
        const E p = h[parent];
        if (!(p.f > v.f)) break;          // comp(first + parent, value)
        h[hole] = p;
        hole = parent;
    }
    h[hole] = v;
}

// Removes and returns the top; `len` is the size before the call (len >= 1).
template <class E, class H>
PDMPC_HD E heap_pop_serial(H h, int len) {
    const E top = h[0];
    const int n = len - 1;               // size after the pop; h[n] is re-inserted
    if (n > 0) {
        const E v = h[n];
        int hole = 0;
        for (;;) {
            const int l = 2 * hole + 1, r = l + 1;
            E c;
            int ci;
            if (r < n) {                 // both children: right unless f_right > f_left
                const E el = h[l], er = h[r];
                if (er.f > el.f) { c = el; ci = l; }
                else { c = er; ci = r; }
            } else if (l < n) {          // n even, single left child at n - 1
                c = h[l];
                ci = l;
            } else {
                break;
            }
            if (c.f > v.f) break;        // __push_heap would move this entry (and all below) back
            h[hole] = c;
            hole = ci;
        }
        h[hole] = v;
    }
    return top;
}

}  // namespace pdmpc
