// pdmpc_kernels.cuh — device code of the B200-native MPA graph search.
//
// One CTA (one warp) runs one search (vehicle x permutation x scenario) and
// reproduces the reference's best-first loop exactly:
//   GraphSearch.do_graph_search      hlc/optimizer/graph_search/GraphSearch.m:23-109
//   eval_edge_exact                  GraphSearch.m:111-196
//   expand_node                      hlc/optimizer/graph_search/expand_node.m:1-91
//   priority queue                   .../priority_queue/priority_queue_interface_mex.cpp:19-108
//                                    == libstdc++ __push_heap / __adjust_heap on (id, f)
//   SAT / lanelet / InterX checkers  intersect_sat.m, intersect_lanelet_boundary.m, InterX.m
//
// Bit-exactness rules: IEEE double everywhere, compiled with --fmad=false (no
// contraction), the same operation order as the reference's expressions, and
// the sin/cos algorithm of DESIGN.md §sincos.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pdmpc_b200.h"

namespace pdmpc {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kAreaStride = PDMPC_AREA_STRIDE;
constexpr int kMaxHp = PDMPC_MAX_HP;

// ---- device views -----------------------------------------------------------
struct MpaDev {
    int nT, Hp, nE;
    const int *succ_ptr;        // [Hp*nT + 1] successors of (step k, trim t) at (k-1)*nT + (t-1)
    const int16_t *succ_trim;   // 1-based end trim, ascending (expand_node.m:18)
    const int16_t *succ_edge;   // edge index of (t, succ_trim)
    const int16_t *edge_of;     // [nT*nT] -> edge or -1
    const double *edge_dx, *edge_dy, *edge_dyaw;
    const int *area_npts;       // [nE*3]
    const double *area_x, *area_y;  // [nE*3*8]
};

struct BatchDev {
    int n, checker;
    double dt;
    const double *x0, *y0, *yaw0;
    const int *trim0;
    const double *ref_x, *ref_y, *v_ref;
    // raw CSR (SAT checker)
    const int *slot_ptr, *poly_ptr;
    const double *vert_x, *vert_y;
    // NaN-separated polylines (InterX checker), vectorize_all_obstacles.m:36-63:
    // vertex v of polygon p sits at v + p, a NaN column at poly_ptr[p+1] + p.
    const double *pl_x, *pl_y;
    // lanelet bounds: raw (SAT) and [left, NaN, right, NaN] per search (InterX),
    // side s of search i at lane_ptr[2i+s] + 2i + s.
    const int *lane_ptr;
    const double *lane_x, *lane_y;
    const double *ll_x, *ll_y;
};

struct OutDev {
    int *status;
    uint8_t *is_exhausted;
    int *n_expanded, *n_pops;
    unsigned long long *pop_hash;
    int *trims, *tree_path;
    double *y_predicted, *g_path, *h_path;
    int *shape_npts;
    double *shape_x, *shape_y;
    unsigned long long *counters;  // [0] pops, [1] nodes, [2] obstacle columns tested
};

// Node arena + heap overflow of one resident CTA ("slot"), all in HBM.
struct __align__(16) NodeA {  // 32 B: pose + cost to come
    double x, y, yaw, g;
};
struct __align__(16) NodeB {  // 16 B
    double h;
    unsigned parent;          // 1-based, 0 for the root (Tree.m:18)
    unsigned short trim;      // 1-based
    unsigned short k;
};
struct ArenaDev {
    NodeA *a;                 // [slots * cap]; index 0 of each slot unused (ids are 1-based)
    NodeB *b;
    double *heap_f;           // [slots * cap] overflow of the shared-memory heap top
    unsigned *heap_id;
    int cap;                  // nodes per slot
};

struct TraceDev {
    int search;               // -1: off
    long long *ids;
    long long cap;
    long long *n;
};

// ---- sin / cos (oracle/pdmpc_oracle.c holds the CPU twin) --
__device__ __forceinline__ void sincos_ref(double x, double &s, double &c) {
    const double TWO_OVER_PI = 0x1.45f306dc9c883p-1;
    const double P1 = 0x1.921fb54400000p+0, P2 = 0x1.0b4611a600000p-34, P3 = 0x1.3198a2e037073p-69;
    const double S[10] = {-0x1.5555555555555p-3, 0x1.1111111111111p-7, -0x1.a01a01a01a01ap-13,
                          0x1.71de3a556c734p-19, -0x1.ae64567f544e4p-26, 0x1.6124613a86d09p-33,
                          -0x1.ae7f3e733b81fp-41, 0x1.952c77030ad4ap-49, -0x1.2f49b46814157p-57,
                          0x1.71b8ef6dcf572p-66};
    const double C[10] = {-0x1.0000000000000p-1, 0x1.5555555555555p-5, -0x1.6c16c16c16c17p-10,
                          0x1.a01a01a01a01ap-16, -0x1.27e4fb7789f5cp-22, 0x1.1eed8eff8d898p-29,
                          -0x1.93974a8c07c9dp-37, 0x1.ae7f3e733b81fp-45, -0x1.6827863b97d97p-53,
                          0x1.e542ba4020225p-62};
    double n = rint(x * TWO_OVER_PI);
    double r = ((x - n * P1) - n * P2) - n * P3;
    double z = r * r;
    double ps = S[9], pc = C[9];
#pragma unroll
    for (int i = 8; i >= 0; --i) {
        ps = ps * z + S[i];
        pc = pc * z + C[i];
    }
    double sr = r + (r * z) * ps;
    double cr = 1.0 + z * pc;
    long long q = (long long)n & 3LL;
    if (q == 0) { s = sr; c = cr; }
    else if (q == 1) { s = cr; c = -sr; }
    else if (q == 2) { s = -sr; c = -cr; }
    else { s = -cr; c = sr; }
}

// ---- priority queue: libstdc++ heap on (f, id), min on f, ties by heap mechanics.
// Entries [0, HS) live in shared memory (the top levels of the array heap = the
// hottest ones), the rest in the slot's HBM overflow.  Run by lane 0 only.
template <int HS>
struct Heap {
    double *sf;
    unsigned *sid;
    double *gf;
    unsigned *gid;
    int len;

    __device__ __forceinline__ double f(int i) const { return i < HS ? sf[i] : gf[i]; }
    __device__ __forceinline__ unsigned id(int i) const { return i < HS ? sid[i] : gid[i]; }
    __device__ __forceinline__ void set(int i, double fv, unsigned iv) {
        if (i < HS) { sf[i] = fv; sid[i] = iv; }
        else { gf[i] = fv; gid[i] = iv; }
    }
    // stl_heap.h:135-147 __push_heap, comp(parent, value) == f(parent) > value
    __device__ __forceinline__ void sift_up(int hole, int top, double fv, unsigned iv) {
        int parent = (hole - 1) / 2;
        while (hole > top) {
            double pf = f(parent);
            if (!(pf > fv)) break;
            set(hole, pf, id(parent));
            hole = parent;
            parent = (hole - 1) / 2;
        }
        set(hole, fv, iv);
    }
    __device__ __forceinline__ void push(double fv, unsigned iv) {  // push_back + push_heap
        int hole = len++;
        sift_up(hole, 0, fv, iv);
    }
    // pop_heap + pop_back; stl_heap.h:224-249 __adjust_heap
    __device__ __forceinline__ unsigned pop() {
        unsigned top_id = id(0);
        if (len > 1) {
            int n = len - 1;
            double vf = f(n);
            unsigned vi = id(n);
            int hole = 0, second = 0;
            while (second < (n - 1) / 2) {
                second = 2 * (second + 1);
                double fr = f(second), fl = f(second - 1);
                if (fr > fl) { second--; fr = fl; }   // comp(right, left) -> take left
                set(hole, fr, id(second));
                hole = second;
            }
            if ((n & 1) == 0 && second == (n - 2) / 2) {
                second = 2 * (second + 1);
                set(hole, f(second - 1), id(second - 1));
                hole = second - 1;
            }
            sift_up(hole, 0, vf, vi);
        }
        --len;
        return top_id;
    }
};

// ---- InterX (InterX.m:63-85,108-110) of a shape with NE segments against the
// NaN-separated polyline points [lo, hi).  Lanes own contiguous runs of
// segments; C1 (shape edge i separates the two obstacle points) is evaluated for
// every pair, C2 only where C1 holds.  The boolean any(C1 & C2) is the same as
// the reference's dense evaluation.
template <int NE>
__device__ __forceinline__ bool interx_range(const double *__restrict__ px, const double *__restrict__ py,
                                             int lo, int hi, const double *shx, const double *shy,
                                             int lane) {
    const int nseg = hi - lo - 1;
    if (nseg < 1) return false;   // InterX.m:48-52 and single-column inputs
    double dx1[NE], dy1[NE], S1[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) {
        dx1[i] = shx[i + 1] - shx[i];
        dy1[i] = shy[i + 1] - shy[i];
        S1[i] = dx1[i] * shy[i] - dy1[i] * shx[i];
    }
    const int per = (nseg + kWarp - 1) / kWarp;
    int j0 = lo + lane * per;
    int j1 = min(j0 + per, lo + nseg);
    bool hit = false;
    if (j0 < j1) {
        double x = __ldg(px + j0), y = __ldg(py + j0);
        double a[NE];
#pragma unroll
        for (int i = 0; i < NE; ++i) a[i] = (dx1[i] * y - dy1[i] * x) - S1[i];
        for (int j = j0; j < j1; ++j) {
            double xn = __ldg(px + j + 1), yn = __ldg(py + j + 1);
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                double an = (dx1[i] * yn - dy1[i] * xn) - S1[i];
                if (a[i] * an < 0) {                       // C1(i,j)
                    double dx2 = xn - x, dy2 = yn - y;
                    double S2 = dx2 * y - dy2 * x;
                    double b0 = (shy[i] * dx2 - shx[i] * dy2) - S2;
                    double b1 = (shy[i + 1] * dx2 - shx[i + 1] * dy2) - S2;
                    if (b0 * b1 < 0) hit = true;           // C2(i,j)
                }
                a[i] = an;
            }
            x = xn;
            y = yn;
        }
    }
    return __any_sync(kFull, hit);
}

__device__ __forceinline__ bool interx_dispatch(int ns, const double *px, const double *py, int lo, int hi,
                                                const double *shx, const double *shy, int lane) {
    switch (ns) {   // ns points -> ns-1 segments; maneuver areas have 5, 6 or 7 points
    case 5: return interx_range<4>(px, py, lo, hi, shx, shy, lane);
    case 6: return interx_range<5>(px, py, lo, hi, shx, shy, lane);
    case 7: return interx_range<6>(px, py, lo, hi, shx, shy, lane);
    default: {
        bool hit = false;   // generic (never taken with MPA areas): one shape segment at a time
        for (int i = 0; i + 1 < ns && !hit; ++i) hit = interx_range<1>(px, py, lo, hi, shx + i, shy + i, lane);
        return hit;
    }
    }
}

// ---- SAT (intersect_sat.m:1-42): polygon 1 in shared memory, polygon 2 in
// global memory.  Lanes own axes (edges of both polygons incl. the closing one).
__device__ __forceinline__ bool sat_collide(const double *x1, const double *y1, int n1,
                                            const double *__restrict__ x2, const double *__restrict__ y2,
                                            int n2, int lane) {
    bool sep = false;
    for (int e = lane; e < n1 + n2; e += kWarp) {
        double ex, ey;
        if (e < n1) {
            int e1 = (e + 1 == n1) ? 0 : e + 1;
            ex = x1[e1] - x1[e];
            ey = y1[e1] - y1[e];
        } else {
            int f = e - n1, f1 = (f + 1 == n2) ? 0 : f + 1;
            ex = __ldg(x2 + f1) - __ldg(x2 + f);
            ey = __ldg(y2 + f1) - __ldg(y2 + f);
        }
        double ax = -ey, ay = ex;
        double nrm = sqrt(ax * ax + ay * ay);
        double nx = ax / nrm, ny = ay / nrm;   // zero edge -> NaN axis -> never separates
        double mn1 = nan(""), mx1 = nan(""), mn2 = nan(""), mx2 = nan("");
        for (int v = 0; v < n1; ++v) {
            double d = nx * x1[v] + ny * y1[v];
            mn1 = fmin(mn1, d);
            mx1 = fmax(mx1, d);
        }
        for (int v = 0; v < n2; ++v) {
            double d = nx * __ldg(x2 + v) + ny * __ldg(y2 + v);
            mn2 = fmin(mn2, d);
            mx2 = fmax(mx2, d);
        }
        // d1/d2 of intersect_a_b; which polygon owns the axis only swaps them
        if ((mn1 - mx2 > 0) || (mn2 - mx1 > 0)) sep = true;
    }
    return !__any_sync(kFull, sep);
}

// ---- intersect_lanelet_boundary.m:1-56 for one side; lanes own boundary segments.
__device__ __forceinline__ bool lanelet_side_sat(const double *sx, const double *sy, int ns,
                                                 const double *__restrict__ bx, const double *__restrict__ by,
                                                 int nb, int lane) {
    if (nb < 2) return false;
    double max_x = sx[0], min_x = sx[0], max_y = sy[0], min_y = sy[0];
    for (int i = 1; i < ns; ++i) {
        max_x = fmax(max_x, sx[i]);
        min_x = fmin(min_x, sx[i]);
        max_y = fmax(max_y, sy[i]);
        min_y = fmin(min_y, sy[i]);
    }
    bool hit = false;
    for (int n = lane; n + 1 < nb; n += kWarp) {
        double ax = __ldg(bx + n), bx2 = __ldg(bx + n + 1), ay = __ldg(by + n), by2 = __ldg(by + n + 1);
        if ((max_x < ax && max_x < bx2) || (min_x > ax && min_x > bx2) || (max_y < ay && max_y < by2) ||
            (min_y > ay && min_y > by2))
            continue;
        // intersect_sat(shape, segment): ns shape axes + the segment's two (opposite) axes
        double qx[2] = {ax, bx2}, qy[2] = {ay, by2};
        bool sep = false;
        for (int e = 0; e < ns + 2 && !sep; ++e) {
            double ex, ey;
            if (e < ns) {
                int e1 = (e + 1 == ns) ? 0 : e + 1;
                ex = sx[e1] - sx[e];
                ey = sy[e1] - sy[e];
            } else {
                int f = e - ns;
                ex = qx[1 - f] - qx[f];
                ey = qy[1 - f] - qy[f];
            }
            double nxx = -ey, nyy = ex;
            double nrm = sqrt(nxx * nxx + nyy * nyy);
            double nx = nxx / nrm, ny = nyy / nrm;
            double mn1 = nan(""), mx1 = nan("");
            for (int v = 0; v < ns; ++v) {
                double d = nx * sx[v] + ny * sy[v];
                mn1 = fmin(mn1, d);
                mx1 = fmax(mx1, d);
            }
            double d0 = nx * qx[0] + ny * qy[0], d1 = nx * qx[1] + ny * qy[1];
            double mn2 = fmin(d0, d1), mx2 = fmax(d0, d1);
            if ((mn1 - mx2 > 0) || (mn2 - mx1 > 0)) sep = true;
        }
        if (!sep) hit = true;
    }
    return __any_sync(kFull, hit);
}

__device__ __forceinline__ unsigned long long fnv1a_u32(unsigned long long h, unsigned v) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        h ^= (unsigned long long)((v >> (8 * b)) & 0xffu);
        h *= 0x100000001b3ULL;
    }
    return h;
}

// Rotate/translate one maneuver area into smem: GraphSearch.m:158-159.
__device__ __forceinline__ void place_point(const MpaDev &m, int edge, int kind, int i, double c, double s,
                                            double px, double py, double &ox, double &oy) {
    const int base = (edge * 3 + kind) * kAreaStride + i;
    double ax = __ldg(m.area_x + base), ay = __ldg(m.area_y + base);
    ox = c * ax - s * ay + px;
    oy = s * ax + c * ay + py;
}

// ============================================================================
// The search kernel.  Persistent: CTA `blockIdx.x` owns arena slot blockIdx.x
// and pulls search indices from a global counter until the batch is drained.
template <int HS>
__global__ void __launch_bounds__(kWarp) search_kernel(MpaDev m, BatchDev b, OutDev o, ArenaDev ar,
                                                       unsigned *work_counter, TraceDev tr) {
    __shared__ double s_heap_f[HS];
    __shared__ unsigned s_heap_id[HS];
    __shared__ double s_refx[kMaxHp], s_refy[kMaxHp], s_vref[kMaxHp];
    __shared__ double s_shx[kAreaStride], s_shy[kAreaStride];   // shape (normal offset)
    __shared__ double s_bhx[kAreaStride], s_bhy[kAreaStride];   // boundary-check shape
    __shared__ double s_f[kWarp];                               // children's f, push order

    const int lane = threadIdx.x;
    const int Hp = m.Hp, nT = m.nT;
    const size_t slot_base = (size_t)blockIdx.x * (size_t)ar.cap;
    NodeA *__restrict__ na = ar.a + slot_base;
    NodeB *__restrict__ nb = ar.b + slot_base;
    Heap<HS> heap;
    heap.sf = s_heap_f;
    heap.sid = s_heap_id;
    heap.gf = ar.heap_f + slot_base;
    heap.gid = ar.heap_id + slot_base;

    for (;;) {
        unsigned si_u = 0;
        if (lane == 0) si_u = atomicAdd(work_counter, 1u);
        si_u = __shfl_sync(kFull, si_u, 0);
        if (si_u >= (unsigned)b.n) break;
        const int si = (int)si_u;

        // ---- per-search set-up ------------------------------------------------
        __syncwarp();
        if (lane < Hp) {
            s_refx[lane] = __ldg(b.ref_x + (size_t)si * Hp + lane);
            s_refy[lane] = __ldg(b.ref_y + (size_t)si * Hp + lane);
            s_vref[lane] = __ldg(b.v_ref + (size_t)si * Hp + lane);
        }
        const int *slot = b.slot_ptr + (size_t)si * (Hp + 1);
        const int trim0 = __ldg(b.trim0 + si);
        if (lane == 0) {   // root: GraphSearch.m:34-46
            NodeA ra;
            ra.x = __ldg(b.x0 + si); ra.y = __ldg(b.y0 + si); ra.yaw = __ldg(b.yaw0 + si); ra.g = 0.0;
            NodeB rb;
            rb.h = 0.0; rb.parent = 0; rb.trim = (unsigned short)trim0; rb.k = 0;
            na[1] = ra;
            nb[1] = rb;
            heap.len = 0;
            heap.push(0.0, 1u);
        }
        __syncwarp();
        // static-obstacle polyline range and lanelet polyline range (InterX layout)
        const int sp0 = __ldg(slot + 0), sp1 = __ldg(slot + 1);
        const int st_lo = __ldg(b.poly_ptr + sp0) + sp0, st_hi = __ldg(b.poly_ptr + sp1) + sp1;
        const int lp0 = __ldg(b.lane_ptr + 2 * si), lp1 = __ldg(b.lane_ptr + 2 * si + 1),
                  lp2 = __ldg(b.lane_ptr + 2 * si + 2);
        const int ll_lo = lp0 + 2 * si, ll_hi = lp2 + 2 * si + 2;

        int n_nodes = 1, n_pops = 0;
        unsigned long long hash = 0xcbf29ce484222325ULL, cols = 0;
        int status = PDMPC_OK;
        bool exhausted = false;
        unsigned goal = 0;
        const bool tracing = (tr.search == si);

        // ---- best-first loop: GraphSearch.m:53-107 ------------------------------
        for (;;) {
            unsigned id = 0;
            if (lane == 0) {
                if (heap.len > 0) id = heap.pop();   // 0 == the reference's -1 (empty)
            }
            id = __shfl_sync(kFull, id, 0);
            if (id == 0) { exhausted = true; break; }   // :57-61
            ++n_pops;
            hash = fnv1a_u32(hash, id);
            if (tracing && lane == 0) {
                if (n_pops <= tr.cap) tr.ids[n_pops - 1] = (long long)id;
                *tr.n = n_pops;
            }

            const NodeB cb = nb[id];
            const NodeA ca = na[id];
            const unsigned par = cb.parent;
            const int cK = cb.k;
            bool valid = true;
            if (par != 0) {   // eval_edge_exact :137-192 (root is valid unchecked)
                const NodeA pa = na[par];
                const int t1 = nb[par].trim, t2 = cb.trim;
                const int edge = __ldg(m.edge_of + (t1 - 1) * nT + (t2 - 1));
                double s, c;
                sincos_ref(pa.yaw, s, c);   // :155-156
                const int ns = __ldg(m.area_npts + edge * 3 + PDMPC_AREA_NORMAL);
                const int bkind = (cK == Hp) ? PDMPC_AREA_LARGE_OFFSET : PDMPC_AREA_WITHOUT_OFFSET;  // :166-174
                const int nbs = __ldg(m.area_npts + edge * 3 + bkind);
                __syncwarp();
                if (lane < ns) place_point(m, edge, PDMPC_AREA_NORMAL, lane, c, s, pa.x, pa.y, s_shx[lane], s_shy[lane]);
                else if (lane >= 8 && lane - 8 < nbs)
                    place_point(m, edge, bkind, lane - 8, c, s, pa.x, pa.y, s_bhx[lane - 8], s_bhy[lane - 8]);
                __syncwarp();

                const int dp0 = __ldg(slot + cK), dp1 = __ldg(slot + cK + 1);
                if (b.checker == PDMPC_CHECKER_INTERX) {
                    // are_constraints_satisfied_interx.m:17,34 (no HDVs)
                    const int dy_lo = __ldg(b.poly_ptr + dp0) + dp0, dy_hi = __ldg(b.poly_ptr + dp1) + dp1;
                    cols += (unsigned long long)((st_hi - st_lo) + (dy_hi - dy_lo) + (ll_hi - ll_lo));
                    if (interx_dispatch(ns, b.pl_x, b.pl_y, st_lo, st_hi, s_shx, s_shy, lane)) valid = false;
                    else if (interx_dispatch(ns, b.pl_x, b.pl_y, dy_lo, dy_hi, s_shx, s_shy, lane)) valid = false;
                    else if (interx_dispatch(nbs, b.ll_x, b.ll_y, ll_lo, ll_hi, s_bhx, s_bhy, lane)) valid = false;
                } else {
                    // are_constraints_satisfied_sat.m:15-53 (nV == 1; HDV block unreachable)
                    for (int pass = 0; pass < 2 && valid; ++pass) {
                        const int q0 = pass == 0 ? sp0 : dp0, q1 = pass == 0 ? sp1 : dp1;
                        for (int p = q0; p < q1 && valid; ++p) {
                            const int v0 = __ldg(b.poly_ptr + p), v1 = __ldg(b.poly_ptr + p + 1);
                            cols += (unsigned long long)(v1 - v0);
                            if (sat_collide(s_shx, s_shy, ns, b.vert_x + v0, b.vert_y + v0, v1 - v0, lane))
                                valid = false;
                        }
                    }
                    if (valid) {
                        cols += (unsigned long long)(lp2 - lp0);
                        if (lanelet_side_sat(s_bhx, s_bhy, nbs, b.lane_x + lp0, b.lane_y + lp0, lp1 - lp0, lane))
                            valid = false;
                        else if (lanelet_side_sat(s_bhx, s_bhy, nbs, b.lane_x + lp1, b.lane_y + lp1, lp2 - lp1, lane))
                            valid = false;
                    }
                }
            }
            if (!valid) continue;                      // :75-77
            if (cK == Hp) { goal = id; break; }        // :81-90

            // ---- expand_node.m:1-91 (nV == 1) ----------------------------------
            const int k_exp = cK + 1;
            const int sbase = __ldg(m.succ_ptr + (k_exp - 1) * nT + (cb.trim - 1));
            const int nchild = __ldg(m.succ_ptr + (k_exp - 1) * nT + (cb.trim - 1) + 1) - sbase;
            if (n_nodes + nchild >= ar.cap) { status = PDMPC_ERR_CAPACITY; break; }
            double s, c;
            sincos_ref(ca.yaw, s, c);                   // :50-51
            const int to_go = Hp - k_exp;               // :37
            for (int c0 = 0; c0 < nchild; c0 += kWarp) {
                const int ci = c0 + lane;
                const unsigned nid = (unsigned)(n_nodes + 1 + ci);
                if (ci < nchild) {
                    const int t2 = __ldg(m.succ_trim + sbase + ci);
                    const int edge = __ldg(m.succ_edge + sbase + ci);
                    const double dx = __ldg(m.edge_dx + edge), dy = __ldg(m.edge_dy + edge),
                                 dyaw = __ldg(m.edge_dyaw + edge);
                    NodeA ea;
                    ea.x = c * dx - s * dy + ca.x;      // :53
                    ea.y = s * dx + c * dy + ca.y;      // :54
                    ea.yaw = ca.yaw + dyaw;             // :55
                    const double ddx = ea.x - s_refx[k_exp - 1], ddy = ea.y - s_refy[k_exp - 1];
                    const double nrm = sqrt(ddx * ddx + ddy * ddy);
                    ea.g = ca.g + nrm * nrm;            // :61
                    double eh = 0.0, d_max = 0.0;       // :66-73
                    for (int it = 1; it <= to_go; ++it) {
                        d_max = d_max + b.dt * s_vref[k_exp + it - 1];
                        const double hx = ea.x - s_refx[k_exp + it - 1], hy = ea.y - s_refy[k_exp + it - 1];
                        const double hn = sqrt(hx * hx + hy * hy);
                        const double mm = fmax(0.0, hn - d_max);
                        eh = eh + mm * mm;
                    }
                    NodeB eb;
                    eb.h = eh; eb.parent = id; eb.trim = (unsigned short)t2; eb.k = (unsigned short)k_exp;
                    na[nid] = ea;                       // Tree.m:54-70 add_nodes
                    nb[nid] = eb;
                    s_f[lane] = ea.g + eh;              // GraphSearch.m:102 (weights 1)
                }
                __syncwarp();
                if (lane == 0) {                        // :104, one push per child in order
                    const int cnt = min(kWarp, nchild - c0);
                    for (int q = 0; q < cnt; ++q) heap.push(s_f[q], (unsigned)(n_nodes + 1 + c0 + q));
                }
                __syncwarp();
            }
            n_nodes += nchild;
        }

        // ---- results: GraphSearch.m:58-60 / :82-89 ------------------------------
        __syncwarp();
        if (status != PDMPC_OK) exhausted = true;   // outputs take the "no plan" defaults
        unsigned path_id = 0;     // lane d holds path[d], d = 0..Hp
        {
            unsigned cur = goal;
            for (int d = Hp; d >= 0; --d) {           // Tree.m:44-52 path_to_root, flipped
                if (lane == d) path_id = cur;
                if (!exhausted && d > 0) cur = nb[cur].parent;
            }
        }
        if (lane == 0) {
            o.status[si] = status;
            if (o.is_exhausted) o.is_exhausted[si] = exhausted ? 1 : 0;
            if (o.n_expanded) o.n_expanded[si] = n_nodes;
            if (o.n_pops) o.n_pops[si] = n_pops;
            if (o.pop_hash) o.pop_hash[si] = hash;
            atomicAdd(o.counters + 0, (unsigned long long)n_pops);
            atomicAdd(o.counters + 1, (unsigned long long)n_nodes);
            atomicAdd(o.counters + 2, cols);
        }
        const double qnan = nan("");
        NodeA pa_l = {qnan, qnan, qnan, qnan};
        NodeB pb_l;
        pb_l.h = qnan; pb_l.parent = 0; pb_l.trim = 0; pb_l.k = 0;
        if (lane <= Hp && !exhausted) { pa_l = na[path_id]; pb_l = nb[path_id]; }
        if (lane <= Hp) {
            const size_t oo = (size_t)si * (Hp + 1) + lane;
            if (o.trims) o.trims[oo] = exhausted ? (lane == 0 ? trim0 : 0) : (int)pb_l.trim;
            if (o.tree_path) o.tree_path[oo] = exhausted ? 0 : (int)path_id;
            if (o.g_path) o.g_path[oo] = pa_l.g;
            if (o.h_path) o.h_path[oo] = pb_l.h;
            if (lane >= 1 && o.y_predicted) {          // return_path_to.m:11-25
                const size_t oy = ((size_t)si * Hp + (lane - 1)) * 3;
                o.y_predicted[oy + 0] = pa_l.x;
                o.y_predicted[oy + 1] = pa_l.y;
                o.y_predicted[oy + 2] = pa_l.yaw;
            }
        }
        if (o.shape_npts) {                            // return_path_area.m:5-7
            // lane d (1..Hp) needs its parent's pose/trim = lane d-1's values
            const double ppx = __shfl_up_sync(kFull, pa_l.x, 1), ppy = __shfl_up_sync(kFull, pa_l.y, 1),
                         ppyaw = __shfl_up_sync(kFull, pa_l.yaw, 1);
            const int ptrim = __shfl_up_sync(kFull, (int)pb_l.trim, 1);
            int edge = 0, ns = 0;
            double s = 0.0, c = 0.0;
            if (lane >= 1 && lane <= Hp && !exhausted) {
                edge = __ldg(m.edge_of + (ptrim - 1) * nT + ((int)pb_l.trim - 1));
                ns = __ldg(m.area_npts + edge * 3 + PDMPC_AREA_NORMAL);
                sincos_ref(ppyaw, s, c);
            }
            if (lane >= 1 && lane <= Hp) {
                const size_t os = (size_t)si * Hp + (lane - 1);
                o.shape_npts[os] = ns;
                if (o.shape_x && o.shape_y) {
                    for (int i = 0; i < kAreaStride; ++i) {
                        double ox = 0.0, oy = 0.0;
                        if (i < ns) place_point(m, edge, PDMPC_AREA_NORMAL, i, c, s, ppx, ppy, ox, oy);
                        o.shape_x[os * kAreaStride + i] = ox;
                        o.shape_y[os * kAreaStride + i] = oy;
                    }
                }
            }
        }
    }
}

// ---- staging kernels --------------------------------------------------------
// vectorize_all_obstacles.m:36-63: copy polygon p to [poly_ptr[p] + p, ...) and
// append the [NaN; NaN] column.  One thread per polygon.
__global__ void build_polyline_kernel(int n_polys, const int *__restrict__ poly_ptr,
                                      const double *__restrict__ vx, const double *__restrict__ vy,
                                      double *__restrict__ px, double *__restrict__ py) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_polys) return;
    const int v0 = poly_ptr[p], v1 = poly_ptr[p + 1];
    for (int v = v0; v < v1; ++v) {
        px[v + p] = vx[v];
        py[v + p] = vy[v];
    }
    px[v1 + p] = nan("");
    py[v1 + p] = nan("");
}

}  // namespace pdmpc
