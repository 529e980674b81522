// pdmpc_heap_split.cuh — the reference's priority queue as the search kernels run it.
//
//   priority_queue_interface_mex.cpp:19-31,62-101 == std::priority_queue<(id, f)> with
//   comp(a, b) = a.f > b.f, i.e. libstdc++ __push_heap (stl_heap.h:135-147) and
//   __adjust_heap (:224-249) on the array — tie order is defined by these mechanics.
//
// Layout: the first `hs` entries live in shared memory, split into a cost array and a
// payload array; the cost of entry i sits at sf[i + 1], so the two children of any entry
// (2i+1, 2i+2) are ONE aligned 16-byte load and the four grandchildren two more.  Entries
// >= hs spill to the HBM arena (interleaved HEnt).
//
// pop(): the hole walks from the root to a leaf along the smaller children, exactly as
// __adjust_heap does, executed uniformly by the warp (every lane takes the same steps,
// one aligned pair load + one compare per level).  The search is bound by the NUMBER of
// dependent warp instructions (~5.5 cycles each, tools/microbench/heap_pop.cu), so the walk is
// kept minimal — multi-level look-ahead variants were measured and were no faster — and the
// path (lane l <-> level l+1) is reconstructed from the final hole instead of being recorded.  The moves are then done in parallel:
// along the path costs are non-decreasing, so __push_heap's upward pass stops at the deepest
// path entry with f <= v.f: entries above it move up one level, v lands below it, the rest
// stays — the array equals the sequential algorithm's after every operation.
// Included by pdmpc_kernels.cuh (after HEnt / kWarp are defined); do not include directly.
#pragma once

namespace pdmpc {

// Shared-memory accesses by 32-bit shared-window address.  The address of the arrays is taken
// ONCE per kernel (opaque to the compiler, see shared_base_once): with ordinary pointers nvcc
// re-derives the window base (S2R SR_CgaCtaId + LEA, a slow special-register read) in front of
// most accesses of the dependent pop chain.
__device__ __forceinline__ unsigned shared_base_once(const void *p) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p), b;
    asm volatile("mov.u32 %0, %1;" : "=r"(b) : "r"(a));   // not rematerialisable
    return b;
}
__device__ __forceinline__ double lds_f64(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double2 lds_f64x2(unsigned a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long lds_u64(unsigned a) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_u64(unsigned a, unsigned long long v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}

struct HeapSplit {
    unsigned sf;                // shared address of the cost array [hs + 2], 16-byte aligned
    unsigned sw;                // shared address of the payload array [hs]
    HEnt *gl;                   // arena overflow, indexed by entry
    int hs;                     // entries held in shared memory (even)
    int len;

    __device__ __forceinline__ double f_at(int i) const { return i < hs ? lds_f64(sf + 8u * (unsigned)(i + 1)) : gl[i].f; }
    __device__ __forceinline__ HEnt ld(int i) const {
        HEnt e;
        if (i < hs) { e.f = lds_f64(sf + 8u * (unsigned)(i + 1)); e.w = lds_u64(sw + 8u * (unsigned)i); }
        else e = gl[i];
        return e;
    }
    __device__ __forceinline__ void st(int i, const HEnt &e) {
        if (i < hs) { sts_f64(sf + 8u * (unsigned)(i + 1), e.f); sts_u64(sw + 8u * (unsigned)i, e.w); }
        else gl[i] = e;
    }

    // __push_heap(first, p, 0, v) by the whole warp (same scheme as Heap::sift_up_at)
    __device__ __forceinline__ void sift_up_at(int p, const HEnt &v, int lane) {
        const int D = 31 - __clz(p + 1);               // number of ancestors of position p
        int base = 0, T = 0;
        for (;;) {
            const int a = base + lane + 1;             // this lane's ancestor, `a` levels up
            const bool anc = a <= D;
            HEnt e;
            e.f = 0.0; e.w = 0;
            if (anc) e = ld(((p + 1) >> a) - 1);
            const unsigned gt = __ballot_sync(0xffffffffu, anc && e.f > v.f);
            const int run = (gt == 0xffffffffu) ? kWarp : (__ffs(~gt) - 1);   // leading run of greater parents
            __syncwarp();
            if (lane < run) st(((p + 1) >> (a - 1)) - 1, e);
            T = base + run;
            if (run < kWarp || base + kWarp >= D) break;
            base += kWarp;
            __syncwarp();
        }
        if (lane == 0) st(((p + 1) >> T) - 1, v);
    }

    // Is the minimum unique (strictly below both children of the root)?  When it is, WHICH entry a
    // pop returns does not depend on the heap's tie mechanics.
    __device__ __forceinline__ bool min_is_unique() const {
        if (len <= 1) return true;
        const double top = f_at(0);
        if (!(top < f_at(1))) return false;
        if (len > 2 && !(top < f_at(2))) return false;
        return true;
    }

    // pq.pop(); caller guarantees len > 0.  LOOKAHEAD: two-level walk (lower latency, more instructions:
    // for a warp that runs alone on its scheduler — the CTA kernel's master)
    template <bool LOOKAHEAD = false>
    __device__ __forceinline__ HEnt pop(int lane) {
        const int n = len - 1;   // heap size after the pop; entry[n] is re-inserted
        len = n;
        if (n < hs) {
            // ---- whole queue in shared memory: no per-access placement tests -------------------
            HEnt top;
            top.f = lds_f64(sf + 8u); top.w = lds_u64(sw);
            if (n > 0) {
                const double vf = lds_f64(sf + 8u * (unsigned)n + 8u);
                const unsigned long long vw = lds_u64(sw + 8u * (unsigned)n);
                // __adjust_heap's walk: the hole follows the smaller child (right unless right.f >
                // left.f) while both children exist; one aligned 16-byte load per level
                int hole = 0, D = 0;
                const int lim = (n - 1) >> 1;
                if (LOOKAHEAD) {
                    // two levels per step: the children pairs of BOTH children are loaded together with the
                    // hole's own pair (three independent 16-byte loads), so the dependent chain per two levels
                    // is one load + one compare + index arithmetic.  Same path as the one-level walk.
                    const int lim2 = (n - 7) >> 2;          // hole <= lim2: all four grandchildren exist
                    while (hole <= lim2) {
                        const unsigned a = sf + 16u * (unsigned)hole + 16u;
                        const double2 p = lds_f64x2(a);
                        const double2 gl = lds_f64x2(a + 16u * (unsigned)hole + 16u);     // pair of child 2h+1
                        const double2 gr = lds_f64x2(a + 16u * (unsigned)hole + 32u);     // pair of child 2h+2
                        const bool left = p.y > p.x;
                        const int child = 2 * hole + 2 - (left ? 1 : 0);
                        const bool left2 = left ? (gl.y > gl.x) : (gr.y > gr.x);
                        hole = 2 * child + 2 - (left2 ? 1 : 0);
                        D += 2;
                    }
                }
                while (hole < lim) {
                    const double2 p = lds_f64x2(sf + 16u * (unsigned)hole + 16u);
                    hole = 2 * hole + 2 - (p.y > p.x ? 1 : 0);
                    ++D;
                }
                if ((n & 1) == 0 && hole == ((n - 2) >> 1)) {   // single (left) child at n-1
                    hole = n - 1;
                    ++D;
                }
                // the path is determined by its last entry: level l+1 (lane l) is the ancestor
                // D-1-l levels above the final hole
                const int myc = ((hole + 1) >> max(D - 1 - lane, 0)) - 1;
                double ef = 0.0;
                unsigned long long ew = 0;
                if (lane < D) { ef = lds_f64(sf + 8u * (unsigned)myc + 8u); ew = lds_u64(sw + 8u * (unsigned)myc); }
                const unsigned le = __ballot_sync(0xffffffffu, lane < D && !(ef > vf));
                const int M = 32 - __clz(le);              // deepest path entry that stays above v, + 1 (0: none)
                // every lane has read its entry (the ballot is the barrier); entries above M move up
                if (lane < M) {
                    const int par = (myc - 1) >> 1;
                    sts_f64(sf + 8u * (unsigned)par + 8u, ef); sts_u64(sw + 8u * (unsigned)par, ew);
                }
                if (lane == max(M - 1, 0)) {
                    const int at = M ? myc : 0;
                    sts_f64(sf + 8u * (unsigned)at + 8u, vf); sts_u64(sw + 8u * (unsigned)at, vw);
                }
            }
            __syncwarp();
            return top;
        }
        const HEnt top = ld(0);
        {
            const HEnt v = ld(n);
            int hole = 0, D = 0;
            const int lim = (n - 1) / 2;               // hole < lim: both children exist
            while (hole < lim) {
                double fl, fr;
                if (2 * hole + 2 < hs) {
                    const double2 p0 = lds_f64x2(sf + 16u * (unsigned)hole + 16u);
                    fl = p0.x; fr = p0.y;
                } else {
                    fl = f_at(2 * hole + 1); fr = f_at(2 * hole + 2);
                }
                hole = 2 * hole + 1 + !(fr > fl);
                ++D;
            }
            if ((n & 1) == 0 && hole == (n - 2) / 2) {   // single (left) child at n-1
                hole = n - 1;
                ++D;
            }
            const int myc = ((hole + 1) >> max(D - 1 - lane, 0)) - 1;
            HEnt e;
            e.f = 0.0; e.w = 0;
            if (lane < D) e = ld(myc);
            const unsigned le = __ballot_sync(0xffffffffu, lane < D && !(e.f > v.f));
            const int M = 32 - __clz(le);
            __syncwarp();
            if (lane < M) st((myc - 1) >> 1, e);
            if (M == 0) { if (lane == 0) st(0, v); }
            else if (lane == M - 1) st(myc, v);
        }
        __syncwarp();
        return top;
    }

    // pq.push of m <= 32 entries in lane order (lane q holds entry q); same scheme as Heap::push_many
    __device__ __forceinline__ void push_many(const HEnt &mine, int m, int lane) {
        int done = 0;
        while (done < m) {
            const bool act = lane >= done && lane < m;
            const int p = len + (lane - done);
            const int par = (p - 1) >> 1;
            bool need = false;
            const int src = done + max(par - len, 0);
            const double nf = __shfl_sync(0xffffffffu, mine.f, src & (kWarp - 1));
            if (act && p > 0) {
                const double pf = (par < len) ? f_at(par) : nf;
                need = pf > mine.f;
            }
            const unsigned mask = __ballot_sync(0xffffffffu, need);
            const int nfast = mask ? (__ffs(mask) - 1 - done) : (m - done);
            if (act && lane < done + nfast) st(p, mine);
            len += nfast;
            done += nfast;
            __syncwarp();
            if (done < m) {
                HEnt v;
                v.f = __shfl_sync(0xffffffffu, mine.f, done);
                v.w = __shfl_sync(0xffffffffu, mine.w, done);
                sift_up_at(len, v, lane);
                ++len;
                ++done;
                __syncwarp();
            }
        }
    }
};

}  // namespace pdmpc
