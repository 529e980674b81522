// pdmpc_lanes.cuh — lane-per-search variant of the MPA graph search (batch throughput).
//
// The warp-per-search kernel (pdmpc_kernels.cuh) is bound by warp-instruction
// issue: ~1150 warp instructions per pop with most lanes doing uniform work
// (profiles/r01a_search_lat12_ncu.md).  Here every THREAD runs one whole search
// on its own: the reference's loop
//   GraphSearch.do_graph_search   hlc/optimizer/graph_search/GraphSearch.m:23-109
//   eval_edge_exact               GraphSearch.m:111-196
//   expand_node                   hlc/optimizer/graph_search/expand_node.m:1-91
//   InterX                        hlc/optimizer/graph_search/InterX.m:63-85
//   priority queue                priority_queue_interface_mex.cpp:19-108 (pdmpc_heap_serial.h)
// is executed literally, per thread, with the same IEEE-double operation order as
// the warp kernel and the oracle, so a warp instruction does 32 searches' worth of
// useful arithmetic.  The dominant cost becomes what it is in the reference too:
// the (shape edges) x (obstacle columns) products of InterX on the FP64 pipe.
//
// Divergence is bounded by construction:
//   * the loop body is a fixed phase sequence (refill, pop, edge check, expand /
//     finish) with reconvergence after each phase;
//   * maneuver areas are padded to 7 points by repeating the last point (a zero
//     edge: all its InterX products are 0 -> never "< 0"), so the InterX loop is
//     the same 6-edge code for every thread;
//   * searches are handed out in order of their obstacle count, so the threads of
//     a warp run InterX loops of similar length.
// A thread gives a search up (it is re-run from scratch by the warp-per-search
// kernel, whose arena holds the full tree) when it outgrows the thread's small
// arena slot or a pop budget: the tail of long searches is latency-critical and a
// single thread is the slowest way to run it.
//
// InterX-only (every prioritized road-network config); SAT
// batches use the warp kernel.
#pragma once

#include "pdmpc_heap_serial.h"
#include "pdmpc_kernels.cuh"

namespace pdmpc {

constexpr int kLaneEdges = 6;            // InterX edges per (padded) shape
constexpr int kLanePts = kLaneEdges + 1;

struct LaneLimits {
    int pop_limit;        // give up (-> overflow list) after this many pops
    unsigned *ov_count;   // number of searches handed over
    int *ov_list;         // their indices
};

// Heap position i of a thread lives at hp[i + 1]: children (2i+1, 2i+2) then share
// one aligned 32-byte sector.
struct LaneHeapRef {
    HEnt *hp;
    __device__ __forceinline__ HEnt &operator[](int i) const { return hp[i + 1]; }
};

// Place the 7 (padded) points of one maneuver area and derive the per-edge
// constants of InterX.m:63-71 (dx1, dy1, S1 = dx1*y1 - dy1*x1).
__device__ __forceinline__ void lane_place_shape(const double *__restrict__ ax, const double *__restrict__ ay,
                                                 double c, double s, double px, double py,
                                                 double (&vx)[kLanePts], double (&vy)[kLanePts],
                                                 double (&dx1)[kLaneEdges], double (&dy1)[kLaneEdges],
                                                 double (&S1)[kLaneEdges]) {
#pragma unroll
    for (int i = 0; i < kLanePts; ++i) {
        const double lx = ax[i], ly = ay[i];
        vx[i] = c * lx - s * ly + px;     // GraphSearch.m:158-159
        vy[i] = s * lx + c * ly + py;
    }
#pragma unroll
    for (int i = 0; i < kLaneEdges; ++i) {
        dx1[i] = vx[i + 1] - vx[i];
        dy1[i] = vy[i + 1] - vy[i];
        S1[i] = dx1[i] * vy[i] - dy1[i] * vx[i];
    }
}

// InterX of the placed shape against the NaN-separated polyline made of the two
// index ranges [lo0, hi0) ++ [lo1, hi1) of p (vectorize_all_obstacles.m:36-63:
// [static polygons ..., polygons of step k ...], every polygon followed by a NaN
// column, so the seam between the ranges is a NaN pseudo-segment like any other).
__device__ __forceinline__ bool lane_interx(const double2 *__restrict__ p, int lo0, int hi0, int lo1, int hi1,
                                            const double (&vx)[kLanePts], const double (&vy)[kLanePts],
                                            const double (&dx1)[kLaneEdges], const double (&dy1)[kLaneEdges],
                                            const double (&S1)[kLaneEdges]) {
    constexpr int G = 4;                                      // columns per software-pipeline stage
    const int n0 = hi0 - lo0, n = n0 + (hi1 - lo1);
    if (n < 2) return false;                                  // InterX.m:48-52
    const int shift = lo1 - n0 - lo0;                         // index j >= n0 maps to lo1 + (j - n0)
    // Columns past the end are clamped to the last one: a zero-length (or NaN) segment, whose
    // products are 0 (or NaN) and never "< 0" - so the pipeline needs no tail code.
    auto col = [&](int j) -> double2 {
        j = min(j, n - 1);
        return p[lo0 + j + (j >= n0 ? shift : 0)];
    };
    double2 q = col(0);
    double x = q.x, y = q.y;
    double a[kLaneEdges];
#pragma unroll
    for (int i = 0; i < kLaneEdges; ++i) a[i] = (dx1[i] * y - dy1[i] * x) - S1[i];
    bool hit = false;
    double2 cur[G], nxt[G];
#pragma unroll
    for (int u = 0; u < G; ++u) cur[u] = col(1 + u);
#pragma unroll 1
    for (int j = 1; j < n; j += G) {
        // the next stage's columns are requested before this stage's arithmetic starts: each
        // thread streams its own polyline, so nothing else hides the load latency
#pragma unroll
        for (int u = 0; u < G; ++u) nxt[u] = col(j + G + u);
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const double xn = cur[u].x, yn = cur[u].y;
            const double dx2 = xn - x, dy2 = yn - y;
            const double S2 = dx2 * y - dy2 * x;
            double bprev = (vy[0] * dx2 - vx[0] * dy2) - S2;
#pragma unroll
            for (int i = 0; i < kLaneEdges; ++i) {
                const double an = (dx1[i] * yn - dy1[i] * xn) - S1[i];
                const double bn = (vy[i + 1] * dx2 - vx[i + 1] * dy2) - S2;
                hit = hit || ((a[i] * an < 0) && (bprev * bn < 0));   // C1 & C2, InterX.m:72-85
                a[i] = an;
                bprev = bn;
            }
            x = xn;
            y = yn;
        }
#pragma unroll
        for (int u = 0; u < G; ++u) cur[u] = nxt[u];
    }
    return hit;
}

// Results of one finished search (GraphSearch.m:58-60 / :82-89, return_path_to.m,
// return_path_area.m), written by the owning thread.  Same layout as the warp kernel.
__device__ __noinline__ void lane_write_result(const MpaDev &m, const Tables &tb, const OutDev &o, int si, int Hp,
                                               bool exhausted, unsigned goal, int trim0, int n_nodes, int n_pops,
                                               unsigned long long hash, unsigned long long cols,
                                               const NodeA *na, const NodeB *nb, const NodeCS *ncs) {
    o.status[si] = PDMPC_OK;
    if (o.is_exhausted) o.is_exhausted[si] = exhausted ? 1 : 0;
    if (o.n_expanded) o.n_expanded[si] = n_nodes;
    if (o.n_pops) o.n_pops[si] = n_pops;
    if (o.pop_hash) o.pop_hash[si] = hash;
    atomicAdd(o.counters + 0, (unsigned long long)n_pops);
    atomicAdd(o.counters + 1, (unsigned long long)n_nodes);
    atomicAdd(o.counters + 2, cols);
    const double qnan = nan("");
    unsigned cur = goal;
    // walk goal -> root (Tree.m:44-52); the child's record and its parent's are both at hand
    for (int d = Hp; d >= 0; --d) {
        NodeA pa = {qnan, qnan, qnan, qnan};
        NodeB pb;
        pb.h = qnan; pb.parent = 0; pb.edge = 0; pb.trim = 0; pb.k = 0;
        if (!exhausted) { pa = na[cur]; pb = nb[cur]; }
        const size_t oo = (size_t)si * (Hp + 1) + d;
        if (o.trims) o.trims[oo] = exhausted ? (d == 0 ? trim0 : 0) : (int)pb.trim;
        if (o.tree_path) o.tree_path[oo] = exhausted ? 0 : (int)cur;
        if (o.g_path) o.g_path[oo] = pa.g;
        if (o.h_path) o.h_path[oo] = pb.h;
        if (d >= 1) {
            const size_t os = (size_t)si * Hp + (d - 1);
            if (o.y_predicted) {
                o.y_predicted[os * 3 + 0] = pa.x;
                o.y_predicted[os * 3 + 1] = pa.y;
                o.y_predicted[os * 3 + 2] = pa.yaw;
            }
            if (o.shape_npts) {
                int edge = 0, ns = 0;
                NodeA qa = {0.0, 0.0, 0.0, 0.0};
                NodeCS qcs = {0.0, 0.0};
                if (!exhausted) {
                    qa = na[pb.parent];
                    qcs = ncs[pb.parent];      // written when the parent was expanded
                    edge = pb.edge;
                    ns = tb.area_npts[edge * 3 + PDMPC_AREA_NORMAL];
                }
                o.shape_npts[os] = ns;
                if (o.shape_x && o.shape_y) {
                    for (int i = 0; i < kAreaStride; ++i) {
                        double ox = 0.0, oy = 0.0;
                        if (i < ns) place_point(tb, edge, PDMPC_AREA_NORMAL, i, qcs.c, qcs.s, qa.x, qa.y, ox, oy);
                        o.shape_x[os * kAreaStride + i] = ox;
                        o.shape_y[os * kAreaStride + i] = oy;
                    }
                }
            }
            cur = pb.parent;
        }
    }
}

// One thread = one search slot.  Persistent: threads pull searches from a global
// counter until the batch is drained.  `apx/apy` = maneuver areas padded to 7
// points (last point repeated), [nE*3*8].
template <int THREADS, bool SMEM_TABLES>
__global__ void __launch_bounds__(THREADS, 1)
search_lanes_kernel(MpaDev m, const double *__restrict__ g_apx, const double *__restrict__ g_apy, BatchDev b,
                    OutDev o, ArenaDev ar, unsigned *work_counter, LaneLimits lim) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Tables tb;
    const double *apx = g_apx, *apy = g_apy;
    if (SMEM_TABLES) {
        // [mbarrier | succ_ptr | succ_te | edge_d | area_npts | area_x | area_y | apx | apy]
        unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw);
        unsigned char *q = smem_raw + 16;
        int *s_succ_ptr = reinterpret_cast<int *>(q); q += m.bytes_succ_ptr;
        int *s_succ_te = reinterpret_cast<int *>(q); q += m.bytes_succ_te;
        double *s_edge_d = reinterpret_cast<double *>(q); q += m.bytes_edge_d;
        int *s_area_npts = reinterpret_cast<int *>(q); q += m.bytes_area_npts;
        double *s_apx = reinterpret_cast<double *>(q); q += m.bytes_area;
        double *s_apy = reinterpret_cast<double *>(q); q += m.bytes_area;
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_expect_tx(bar, m.bytes_succ_ptr + m.bytes_succ_te + m.bytes_edge_d + m.bytes_area_npts +
                                    2 * m.bytes_area);
            tma_bulk_g2s(s_succ_ptr, m.succ_ptr, m.bytes_succ_ptr, bar);
            tma_bulk_g2s(s_succ_te, m.succ_te, m.bytes_succ_te, bar);
            tma_bulk_g2s(s_edge_d, m.edge_d, m.bytes_edge_d, bar);
            tma_bulk_g2s(s_area_npts, m.area_npts, m.bytes_area_npts, bar);
            tma_bulk_g2s(s_apx, g_apx, m.bytes_area, bar);
            tma_bulk_g2s(s_apy, g_apy, m.bytes_area, bar);
        }
        __syncthreads();
        mbar_wait(bar, 0);
        tb.succ_ptr = s_succ_ptr; tb.succ_te = s_succ_te; tb.edge_d = s_edge_d; tb.area_npts = s_area_npts;
        apx = s_apx; apy = s_apy;
    } else {
        tb.succ_ptr = m.succ_ptr; tb.succ_te = m.succ_te; tb.edge_d = m.edge_d; tb.area_npts = m.area_npts;
    }
    // the unpadded areas are only read when a result is written (rare): keep them in global memory
    tb.area_x = m.area_x; tb.area_y = m.area_y;

    const int Hp = m.Hp, nT = m.nT;
    const size_t slot_base = ((size_t)blockIdx.x * THREADS + threadIdx.x) * (size_t)ar.cap;
    NodeA *__restrict__ na = ar.a + slot_base;
    NodeB *__restrict__ nb = ar.b + slot_base;
    NodeCS *__restrict__ ncs = ar.cs + slot_base;
    LaneHeapRef heap;
    heap.hp = ar.heap + slot_base;
    const int cap = ar.cap - 2;   // heap positions are shifted by one entry

    int si = -1, trim0 = 0, heap_len = 0, n_nodes = 0, n_pops = 0;
    unsigned long long hash = 0, cols = 0;
    const int *rng = nullptr;     // polyline offsets of obstacle slots 0..Hp+1 of the current search
    int llo = 0, lhi = 0;
    const double *refx = nullptr, *refy = nullptr, *vref = nullptr;

    PROF_DECL
    for (;;) {
        PROF_MARK(7);
        // ---- refill ---------------------------------------------------------------
        if (si < 0) {
            const unsigned w = atomicAdd(work_counter, 1u);
            if (w >= (unsigned)b.n) break;
            si = b.order ? __ldg(b.order + w) : (int)w;
            trim0 = __ldg(b.trim0 + si);
            NodeA ra;   // root: GraphSearch.m:34-46
            ra.x = __ldg(b.x0 + si); ra.y = __ldg(b.y0 + si); ra.yaw = __ldg(b.yaw0 + si); ra.g = 0.0;
            NodeB rb;
            rb.h = 0.0; rb.parent = 0; rb.edge = 0xffff; rb.trim = (unsigned char)trim0; rb.k = 0;
            na[1] = ra;
            nb[1] = rb;
            HEnt re;
            re.f = 0.0; re.w = HEnt::pack(1u, 0u, 0u, 0u, (unsigned)trim0);
            heap[0] = re;
            heap_len = 1;
            n_nodes = 1; n_pops = 0;
            hash = 0xcbf29ce484222325ULL; cols = 0;
            rng = b.rng + (size_t)si * (Hp + 2);
            llo = __ldg(b.lane_ptr + 2 * si) + 2 * si;
            lhi = __ldg(b.lane_ptr + 2 * si + 2) + 2 * si + 2;
            refx = b.ref_x + (size_t)si * Hp;
            refy = b.ref_y + (size_t)si * Hp;
            vref = b.v_ref + (size_t)si * Hp;
        }

        PROF_MARK(0);
        // ---- pop: GraphSearch.m:53-61 ----------------------------------------------
        bool finished = false, exhausted = false, give_up = false;
        unsigned goal = 0;
        if (heap_len == 0) {
            finished = true; exhausted = true;
        } else if (n_pops >= lim.pop_limit) {
            give_up = true;
        } else {
            const HEnt top = heap_pop_serial<HEnt>(heap, heap_len);
            --heap_len;
            const unsigned id = top.id(), par = top.pid();
            const int cK = (int)top.k();
            ++n_pops;
            hash = hash_step(hash, id);
            PROF_MARK(1);

            // ---- eval_edge_exact: GraphSearch.m:137-192 ---------------------------
            bool valid = true;
            if (par != 0) {
                const NodeA pa = na[par];
                const NodeCS pcs = ncs[par];          // cos/sin(parent yaw), stored at its expansion
                const int edge = (int)top.edge();
                const int bkind = (cK == Hp) ? PDMPC_AREA_LARGE_OFFSET : PDMPC_AREA_WITHOUT_OFFSET;   // :166-174
                const int st_lo = __ldg(rng + 0), st_hi = __ldg(rng + 1);
                const int dy_lo = __ldg(rng + cK), dy_hi = __ldg(rng + cK + 1);
                cols += (unsigned long long)((st_hi - st_lo) + (dy_hi - dy_lo) + (lhi - llo));
#pragma unroll 1
                for (int pass = 0; pass < 2 && valid; ++pass) {
                    // pass 0: maneuver area vs obstacles of step k; pass 1: boundary-check area vs lanelet bounds
                    // (are_constraints_satisfied_interx.m:17,34; no HDVs)
                    const int kind = pass == 0 ? PDMPC_AREA_NORMAL : bkind;
                    const int base = (edge * 3 + kind) * kAreaStride;
                    double vx[kLanePts], vy[kLanePts], dx1[kLaneEdges], dy1[kLaneEdges], S1[kLaneEdges];
                    lane_place_shape(apx + base, apy + base, pcs.c, pcs.s, pa.x, pa.y, vx, vy, dx1, dy1, S1);
                    const double2 *pp = pass == 0 ? b.pl_xy : b.ll_xy;
                    const int lo0 = pass == 0 ? st_lo : llo, hi0 = pass == 0 ? st_hi : lhi;
                    const int lo1 = pass == 0 ? dy_lo : 0, hi1 = pass == 0 ? dy_hi : 0;
                    if (lane_interx(pp, lo0, hi0, lo1, hi1, vx, vy, dx1, dy1, S1)) valid = false;
                }
            }
            PROF_MARK(3);
            if (valid) {                                              // :75-77 invalid nodes are just skipped
                if (cK == Hp) {                                       // :81-90
                    finished = true; goal = id;
                } else {
                    // ---- expand_node.m:1-91 (nV == 1) -------------------------------
                    const int ctrim = (int)top.trim();
                    const int k_exp = cK + 1;
                    const int sbase = tb.succ_ptr[(k_exp - 1) * nT + (ctrim - 1)];
                    const int nchild = tb.succ_ptr[(k_exp - 1) * nT + (ctrim - 1) + 1] - sbase;
                    if (n_nodes + nchild >= cap) {
                        give_up = true;
                    } else {
                        const NodeA ca = na[id];
                        NodeCS ccs;
                        sincos_ref(ca.yaw, ccs.s, ccs.c);             // :50-51
                        ncs[id] = ccs;                                // for the children's edge checks
                        const double s = ccs.s, c = ccs.c;
                        const int to_go = Hp - k_exp;                 // :37
                        const double rx = __ldg(refx + k_exp - 1), ry = __ldg(refy + k_exp - 1);
#pragma unroll 1
                        for (int ci = 0; ci < nchild; ++ci) {
                            const int te = tb.succ_te[sbase + ci];
                            const int cedge = te >> 8, t2 = (te & 0xff) + 1;
                            const double mdx = tb.edge_d[cedge * 4 + 0], mdy = tb.edge_d[cedge * 4 + 1],
                                         mdyaw = tb.edge_d[cedge * 4 + 2];
                            NodeA ea;
                            ea.x = c * mdx - s * mdy + ca.x;          // :53
                            ea.y = s * mdx + c * mdy + ca.y;          // :54
                            ea.yaw = ca.yaw + mdyaw;                  // :55
                            const double ddx = ea.x - rx, ddy = ea.y - ry;
                            const double nrm = sqrt(ddx * ddx + ddy * ddy);
                            ea.g = ca.g + nrm * nrm;                  // :61
                            double eh = 0.0, d_max = 0.0;             // :66-73
#pragma unroll 1
                            for (int it0 = 1; it0 <= to_go; it0 += 4) {
                                // the square roots of four steps are independent: issue them together
                                double hn[4], dv[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int kk = min(k_exp + it0 + u - 1, Hp - 1);
                                    const double hx = ea.x - __ldg(refx + kk), hy = ea.y - __ldg(refy + kk);
                                    hn[u] = sqrt(hx * hx + hy * hy);
                                    dv[u] = b.dt * __ldg(vref + kk);
                                }
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    if (it0 + u <= to_go) {
                                        d_max = d_max + dv[u];
                                        const double mm = fmax(0.0, hn[u] - d_max);
                                        eh = eh + mm * mm;
                                    }
                                }
                            }
                            const unsigned nid = (unsigned)(n_nodes + 1 + ci);
                            NodeB eb;
                            eb.h = eh; eb.parent = id; eb.edge = (unsigned short)cedge;
                            eb.trim = (unsigned char)t2; eb.k = (unsigned char)k_exp;
                            na[nid] = ea;                             // Tree.m:54-70 add_nodes
                            nb[nid] = eb;
                            HEnt he;
                            he.f = ea.g + eh;                         // GraphSearch.m:102 (weights 1)
                            he.w = HEnt::pack(nid, id, (unsigned)cedge, (unsigned)k_exp, (unsigned)t2);
                            heap_push_serial<HEnt>(heap, heap_len, he);   // :104, one push per child, in order
                            ++heap_len;
                        }
                        n_nodes += nchild;
                    }
                }
            }
        }
        PROF_MARK(4);
        // ---- finish / hand over ----------------------------------------------------
        if (finished) {
            lane_write_result(m, tb, o, si, Hp, exhausted, goal, trim0, n_nodes, n_pops, hash, cols, na, nb, ncs);
            si = -1;
        } else if (give_up) {
            const unsigned slot = atomicAdd(lim.ov_count, 1u);
            lim.ov_list[slot] = si;
            si = -1;
        }
        PROF_MARK(5);
    }
    PROF_FLUSH(o, (threadIdx.x & 31) == 0);
}

// ---- staging: interleaved NaN-separated polylines + per-search slot offsets -------
// vectorize_all_obstacles.m:36-63 as (x, y) pairs: one 16-byte load per column.
__global__ void build_polyline_xy_kernel(int n_polys, const int *__restrict__ poly_ptr,
                                         const double *__restrict__ vx, const double *__restrict__ vy,
                                         double2 *__restrict__ pxy) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_polys) return;
    const int v0 = poly_ptr[p], v1 = poly_ptr[p + 1];
    for (int v = v0; v < v1; ++v) pxy[v + p] = make_double2(vx[v], vy[v]);
    pxy[v1 + p] = make_double2(nan(""), nan(""));
}

// rng[i*(Hp+2) + s] = polyline offset of obstacle slot s of search i (s = Hp+1: end)
__global__ void build_ranges_kernel(int n, int Hp, const int *__restrict__ slot_ptr,
                                    const int *__restrict__ poly_ptr, int *__restrict__ rng) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * (Hp + 2)) return;
    const int i = t / (Hp + 2), s = t % (Hp + 2);
    const int q = slot_ptr[(size_t)i * (Hp + 1) + s];
    rng[t] = poly_ptr[q] + q;
}

}  // namespace pdmpc
