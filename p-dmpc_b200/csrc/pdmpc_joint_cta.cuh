// pdmpc_joint_cta.cuh — centralized (joint) graph search, one CTA per search, FACTORIZED expansion.
//
//   CentralizedController.controller                 hlc/controller/centralized/CentralizedController.m:33-59
//   GraphSearch.do_graph_search / eval_edge_exact    hlc/optimizer/graph_search/GraphSearch.m:23-196
//   expand_node (Cartesian product of successors)    hlc/optimizer/graph_search/expand_node.m:15-75 (cartprod: first
//                                                    vehicle fastest; costs summed over the vehicles in order)
//   are_constraints_satisfied_sat                    .../are_constraints_satisfied_sat.m:15-53
//   priority queue                                   .../priority_queue/priority_queue_interface_mex.cpp:19-108
//
// Same results as joint_search_kernel (pdmpc_joint.cuh, one warp per search, the reference's loop statement by
// statement).  What makes this one 10-30x faster on a single search:
//
// 1. Everything a child needs is a function of (expanded node, vehicle v, successor i_v) or of a PAIR of those —
//    never of the whole combination.  A node with n_v successors per vehicle has prod(n_v) children (144 for two
//    vehicles, 1728 for three) but only sum(n_v) poses / cost terms / placed areas / obstacle and lanelet tests
//    and sum over pairs of n_u * n_v vehicle-vs-vehicle tests:
//      * eval_edge_exact places the maneuver's areas by the PARENT pose (GraphSearch.m:155-160): the shapes of a
//        child depend on the expanded node and the vehicle's own successor only;
//      * the verdict is a conjunction (static obstacles, dynamic obstacles of the step, vehicles u < v, own lanelet
//        boundary; are_constraints_satisfied_sat.m:15-53), so it does not depend on the order of the tests;
//      * g and h are sums of per-vehicle terms added vehicle by vehicle, step by step (expand_node.m:43-75): the
//        terms are computed once per (v, i_v), the child's sums add them in the reference's order (same values,
//        same order of additions -> same bits).
//    So whether a child is valid is known when it is CREATED, at a fraction of the cost of checking it when popped.
// 2. A node record is (g, h, parent, depth, child index): a child's vehicle states are recomputed from the parent's
//    when (if) the child is popped — the same expressions on the same inputs, hence the same bits.  Vehicle states
//    are stored for expanded nodes only.
// 3. Two queue disciplines, as in pdmpc_cta.cuh:
//      exact       every child is pushed (valid or not); a popped invalid child is skipped.  The queue sees the
//                  reference's pushes and pops in the reference's order: identical tie-breaking, full pop_hash.
//      valid-only  (b.hash_valid_only, i.e. pdmpc_set_cta_queue(1)) invalid children are not pushed at all; before
//                  every pop the minimum must be unique, else the search is re-run with the exact discipline.
//                  n_pops is recovered exactly (an invalid node was popped iff its f is below the goal's; an
//                  equal f re-runs).  97 % of the pops of a two-vehicle search are invalid nodes.
//
// One CTA of kJWarps warps per search; phases separated by CTA barriers: pop (warp 0) -> recompute the node ->
// per-(vehicle, successor) items (one warp each) -> pair tests (one thread each) -> children (one thread each,
// ordered compaction) -> pushes (warp 0).
#pragma once

#include "pdmpc_joint.cuh"

namespace pdmpc {

constexpr int kJWarps = 8;
constexpr int kJThreads = kJWarps * kWarp;
constexpr int kJSucc = 16;          // successors of a (trim, step) the factorized tables hold (MPAs here: <= 12)
constexpr int kJCtaHeap = 4096;     // heap entries in shared memory
constexpr int kJCand = 2048;        // children waiting for their push
constexpr int kJPairs = kMaxJoint * (kMaxJoint - 1) / 2;

struct __align__(16) JNode2 {       // 32 B per node
    double g, h;
    unsigned pxs;                   // expanded-record slot of the parent (0: root)
    int k;
    unsigned ci;                    // child index within the parent's expansion | 0x80000000 if invalid
    unsigned pad;
};
struct __align__(16) JXHdr {        // per expanded (popped valid) node, followed by nV JVeh records
    unsigned id, pxs;
    unsigned pad[2];
};
struct JointCtaArena {
    JNode2 *node;                   // [slots * cap]
    unsigned char *xrec;            // [slots * xcap * xstride] expanded records
    HEnt *heap;                     // [slots * cap]
    int cap, xcap, xstride, nV;
};

struct __align__(16) JointCtaSmem {
    double hf[kJCtaHeap + 2];
    unsigned long long hw[kJCtaHeap];
    double refx[kMaxJoint][kMaxHp], refy[kMaxJoint][kMaxHp], vref[kMaxJoint][kMaxHp];
    JVeh cur[kMaxJoint];                                  // vehicle states of the node being expanded
    int nsucc[kMaxJoint], sbase[kMaxJoint];
    double gterm[kMaxJoint][kJSucc];                      // norm(...)^2 of expand_node.m:61
    double hterm[kMaxJoint][kJSucc][kMaxHp];              // max(0, ...)^2 of :66-73, per step to go
    double shx[kMaxJoint][kJSucc][kAreaStride], shy[kMaxJoint][kJSucc][kAreaStride];   // normal-offset areas
    int shn[kMaxJoint][kJSucc];
    double bhx[kJWarps][kAreaStride], bhy[kJWarps][kAreaStride];                       // boundary-check area (per warp)
    unsigned sv[kMaxJoint];                               // bit i: successor i of vehicle v passes its own tests
    unsigned pairbad[kJPairs][kJSucc];                    // [pair(u<v)][i_u] bit i_v: the two areas collide
    double cf[kJCand];                                    // children to push, in child order
    unsigned long long cw[kJCand];
    int ncand, wcount[kJWarps];
    // pop broadcast
    unsigned pop_id;
    int pop_state;                                        // 0 node popped, 1 queue empty, 2 tie
    double pop_f;
    unsigned path_xs[kMaxHp + 1];
    int red[kJWarps];
};

__device__ __forceinline__ int jpair(int u, int v) { return v * (v - 1) / 2 + u; }   // u < v

// intersect_sat.m:1-42 by ONE thread (both polygons in shared memory): the arithmetic of sat_collide, axis by axis
__device__ __forceinline__ bool sat_collide_scalar(const double *x1, const double *y1, int n1, const double *x2,
                                                   const double *y2, int n2) {
    for (int e = 0; e < n1 + n2; ++e) {
        double ex, ey;
        if (e < n1) {
            const int e1 = (e + 1 == n1) ? 0 : e + 1;
            ex = x1[e1] - x1[e];
            ey = y1[e1] - y1[e];
        } else {
            const int f = e - n1, f1 = (f + 1 == n2) ? 0 : f + 1;
            ex = x2[f1] - x2[f];
            ey = y2[f1] - y2[f];
        }
        const double ax = -ey, ay = ex;
        const double nrm = sqrt(ax * ax + ay * ay);
        const double nx = ax / nrm, ny = ay / nrm;
        double mn1 = nan(""), mx1 = nan(""), mn2 = nan(""), mx2 = nan("");
        for (int v = 0; v < n1; ++v) {
            const double d = nx * x1[v] + ny * y1[v];
            mn1 = fmin(mn1, d);
            mx1 = fmax(mx1, d);
        }
        for (int v = 0; v < n2; ++v) {
            const double d = nx * x2[v] + ny * y2[v];
            mn2 = fmin(mn2, d);
            mx2 = fmax(mx2, d);
        }
        if ((mn1 - mx2 > 0) || (mn2 - mx1 > 0)) return false;
    }
    return true;
}

__global__ void __launch_bounds__(kJThreads, 1)
joint_cta_kernel(MpaDev m, BatchDev b, OutDev o, JointCtaArena ar, unsigned *work_counter) {
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char jc_smem_raw[];
    JointCtaSmem &sm = *reinterpret_cast<JointCtaSmem *>(jc_smem_raw);
    Tables tb;
    tb.succ_ptr = m.succ_ptr; tb.succ_te = m.succ_te; tb.edge_d = m.edge_d;
    tb.area_npts = m.area_npts; tb.area_x = m.area_x; tb.area_y = m.area_y;
    const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp;
    Tile<kWarp> t;
    t.shift = 0; t.lane = lane; t.mask = FULL;
    const int Hp = m.Hp, nT = m.nT, nV = ar.nV;
    const int n_joint = b.n / nV;
    JNode2 *__restrict__ nn = ar.node + (size_t)blockIdx.x * ar.cap;
    unsigned char *__restrict__ xr = ar.xrec + (size_t)blockIdx.x * ar.xcap * ar.xstride;
    auto xhdr = [&](unsigned xs) -> JXHdr * { return reinterpret_cast<JXHdr *>(xr + (size_t)xs * ar.xstride); };
    auto xveh = [&](unsigned xs) -> JVeh * { return reinterpret_cast<JVeh *>(xr + (size_t)xs * ar.xstride + sizeof(JXHdr)); };
    HeapSplit heap;
    heap.sf = shared_base_once(sm.hf);
    heap.sw = shared_base_once(sm.hw);
    heap.gl = ar.heap + (size_t)blockIdx.x * ar.cap;
    heap.hs = kJCtaHeap;
    heap.len = 0;
    __shared__ unsigned s_search;
    bool redo_exact = false;
    unsigned su = 0;
    PROF_DECL

    for (;;) {
        // ---- fetch a search (or run the same one again with the exact queue) -----------------------------
        bool exact;
        if (redo_exact) { redo_exact = false; exact = true; }
        else {
            __syncthreads();
            if (tid == 0) s_search = atomicAdd(work_counter, 1u);
            __syncthreads();
            su = s_search;
            exact = b.hash_valid_only == 0;
        }
        if (su >= (unsigned)n_joint) break;
        const int r0 = (int)su * nV;
        for (int i = tid; i < nV * Hp; i += kJThreads) {
            const int v = i / Hp, k = i % Hp;
            sm.refx[v][k] = __ldg(b.ref_x + (size_t)(r0 + v) * Hp + k);
            sm.refy[v][k] = __ldg(b.ref_y + (size_t)(r0 + v) * Hp + k);
            sm.vref[v][k] = __ldg(b.v_ref + (size_t)(r0 + v) * Hp + k);
        }
        if (tid == 0) {   // root: GraphSearch.m:34-46
            JNode2 rn;
            rn.g = 0.0; rn.h = 0.0; rn.pxs = 0; rn.k = 0; rn.ci = 0; rn.pad = 0;
            nn[1] = rn;
            HEnt re;
            re.f = 0.0; re.w = jpack(1u, 0u);
            heap.st(0, re);
            sm.ncand = 0;
        }
        heap.len = 1;   // (every thread keeps the length: warp 0 is the only one that uses the heap)
        const int *slot = b.slot_ptr + (size_t)r0 * (Hp + 1);
        const int sp0 = __ldg(slot + 0), sp1 = __ldg(slot + 1);
        int n_nodes = 1, n_pops = 0, n_x = 0, status = PDMPC_OK;
        unsigned long long hash = 0xcbf29ce484222325ULL, cols = 0;
        bool exhausted = false, tie = false;
        unsigned goal_xs = 0;
        double f_goal = 0.0;
        __syncthreads();

        PROF_MARK(0);   // set-up
        for (;;) {   // GraphSearch.m:53-107
            // ---- pop (warp 0) ----------------------------------------------------------------------------
            if (warp == 0) {
                int st = 0;
                unsigned id = 0;
                double f = 0.0;
                for (;;) {
                    if (heap.len == 0) { st = 1; break; }                       // :57-61
                    if (!exact && !heap.min_is_unique()) { st = 2; break; }     // tie mechanics would matter
                    const HEnt top = heap.pop(lane);
                    id = (unsigned)(top.w & 0xfffffffULL);
                    f = top.f;
                    ++n_pops;
                    if (!b.hash_valid_only) hash = hash_step(hash, id);
                    if (top.w >> 63) continue;                                  // created invalid: :75-77
                    if (b.hash_valid_only) hash = hash_step(hash, id);
                    break;
                }
                if (lane == 0) { sm.pop_state = st; sm.pop_id = id; sm.pop_f = f; }
            }
            __syncthreads();
            PROF_MARK(1);   // pop
            const int pst = sm.pop_state;
            if (pst == 1) { exhausted = true; break; }
            if (pst == 2) { tie = true; break; }
            const unsigned id = sm.pop_id;
            // ---- the node's vehicle states, recomputed from its parent's (thread v < nV) -----------------
            const JNode2 cn = nn[id];
            const int cK = cn.k;
            if (n_x + 1 >= ar.xcap) { status = PDMPC_ERR_CAPACITY; break; }
            const unsigned xs = (unsigned)(++n_x);
            if (tid < nV) {
                JVeh ev;
                if (cn.pxs == 0) {   // root
                    ev.x = __ldg(b.x0 + r0 + tid); ev.y = __ldg(b.y0 + r0 + tid); ev.yaw = __ldg(b.yaw0 + r0 + tid);
                    ev.edge = 0xffff; ev.trim = (unsigned char)__ldg(b.trim0 + r0 + tid);
                } else {
                    const JVeh *pv = xveh(cn.pxs);
                    long long rem = (long long)(cn.ci & 0x7fffffffu);
                    int iv = 0;
                    for (int v = 0; v <= tid; ++v) {   // cartprod: first vehicle fastest
                        const int q = (cK - 1) * nT + ((int)pv[v].trim - 1);
                        const int ns = tb.succ_ptr[q + 1] - tb.succ_ptr[q];
                        iv = (int)(rem % ns);
                        rem /= ns;
                    }
                    const JVeh p = pv[tid];
                    const int te = tb.succ_te[tb.succ_ptr[(cK - 1) * nT + ((int)p.trim - 1)] + iv];
                    const int cedge = te >> 8;
                    const double mdx = tb.edge_d[cedge * 4 + 0], mdy = tb.edge_d[cedge * 4 + 1], mdyaw = tb.edge_d[cedge * 4 + 2];
                    ev.x = p.c * mdx - p.s * mdy + p.x;      // expand_node.m:53
                    ev.y = p.s * mdx + p.c * mdy + p.y;      // :54
                    ev.yaw = p.yaw + mdyaw;                  // :55
                    ev.edge = (unsigned short)cedge; ev.trim = (unsigned char)((te & 0xff) + 1);
                }
                sincos_ref(ev.yaw, ev.s, ev.c);              // :50-51 (of this node's own expansion)
                ev.pad = 0; ev.pad2 = 0;
                xveh(xs)[tid] = ev;
                sm.cur[tid] = ev;
                if (tid == 0) {
                    JXHdr hd;
                    hd.id = id; hd.pxs = cn.pxs; hd.pad[0] = hd.pad[1] = 0;
                    *xhdr(xs) = hd;
                }
            }
            if (cK == Hp) { goal_xs = xs; f_goal = sm.pop_f; break; }   // :81-90 (valid by construction)

            // ---- expand_node.m ---------------------------------------------------------------------------
            const int k_exp = cK + 1;
            const int to_go = Hp - k_exp;                               // :37
            if (tid < nV) {
                const int q = (k_exp - 1) * nT + ((int)sm.cur[tid].trim - 1);
                sm.sbase[tid] = tb.succ_ptr[q];
                sm.nsucc[tid] = tb.succ_ptr[q + 1] - tb.succ_ptr[q];
                sm.sv[tid] = 0u;
            }
            for (int i = tid; i < kJPairs * kJSucc; i += kJThreads) (&sm.pairbad[0][0])[i] = 0u;
            __syncthreads();
            long long total = 1;
            int n_items = 0, ofs[kMaxJoint + 1];
            bool too_wide = false;
            for (int v = 0; v < nV; ++v) {
                ofs[v] = n_items;
                n_items += sm.nsucc[v];
                total *= sm.nsucc[v];
                if (sm.nsucc[v] > kJSucc) too_wide = true;
            }
            ofs[nV] = n_items;
            if (too_wide || (long long)n_nodes + total >= (long long)ar.cap || total >= 0x7fffffffLL) {
                status = PDMPC_ERR_CAPACITY;
                break;
            }
            PROF_MARK(2);   // node state + expansion set-up
            // ---- phase A: one warp per (vehicle, successor) item ----------------------------------------
            const int dp0 = __ldg(slot + k_exp), dp1 = __ldg(slot + k_exp + 1);
            const int bkind = (k_exp == Hp) ? PDMPC_AREA_LARGE_OFFSET : PDMPC_AREA_WITHOUT_OFFSET;   // GraphSearch.m:166-174
            for (int it = warp; it < n_items; it += kJWarps) {
                int v = 0;
                while (it >= ofs[v + 1]) ++v;
                const int i = it - ofs[v];
                const JVeh cv = sm.cur[v];
                const int te = tb.succ_te[sm.sbase[v] + i];
                const int cedge = te >> 8;
                const int ns = tb.area_npts[cedge * 3 + PDMPC_AREA_NORMAL];
                const int nbs = tb.area_npts[cedge * 3 + bkind];
                // areas of the maneuver placed by the expanded node's pose (the child's PARENT pose)
                if (lane < kAreaStride)
                    place_point(tb, cedge, PDMPC_AREA_NORMAL, lane, cv.c, cv.s, cv.x, cv.y, sm.shx[v][i][lane], sm.shy[v][i][lane]);
                else if (lane < 2 * kAreaStride)
                    place_point(tb, cedge, bkind, lane - kAreaStride, cv.c, cv.s, cv.x, cv.y, sm.bhx[warp][lane - kAreaStride],
                                sm.bhy[warp][lane - kAreaStride]);
                else if (lane == 16) {
                    // cost terms of the child of this vehicle (expand_node.m:53-73)
                    const double mdx = tb.edge_d[cedge * 4 + 0], mdy = tb.edge_d[cedge * 4 + 1];
                    const double ex = cv.c * mdx - cv.s * mdy + cv.x;
                    const double ey = cv.s * mdx + cv.c * mdy + cv.y;
                    const double ddx = ex - sm.refx[v][k_exp - 1], ddy = ey - sm.refy[v][k_exp - 1];
                    const double nrm = sqrt(ddx * ddx + ddy * ddy);
                    sm.gterm[v][i] = nrm * nrm;
                    double d_max = 0.0;
                    for (int s2 = 1; s2 <= to_go; ++s2) {
                        d_max = d_max + b.dt * sm.vref[v][k_exp + s2 - 1];
                        const double hx = ex - sm.refx[v][k_exp + s2 - 1], hy = ey - sm.refy[v][k_exp + s2 - 1];
                        const double mm = fmax(0.0, sqrt(hx * hx + hy * hy) - d_max);
                        sm.hterm[v][i][s2 - 1] = mm * mm;
                    }
                    sm.shn[v][i] = ns;
                }
                __syncwarp();
                // are_constraints_satisfied_sat.m:15-35, :46-53 for this vehicle alone
                bool ok = true;
                for (int pass = 0; pass < 2 && ok; ++pass) {
                    const int q0 = pass == 0 ? sp0 : dp0, q1 = pass == 0 ? sp1 : dp1;
                    for (int p = q0; p < q1 && ok; ++p) {
                        const int v0 = __ldg(b.poly_ptr + p), v1 = __ldg(b.poly_ptr + p + 1);
                        cols += (unsigned long long)(v1 - v0);
                        if (sat_collide<kWarp>(sm.shx[v][i], sm.shy[v][i], ns, b.vert_x + v0, b.vert_y + v0, v1 - v0, t)) ok = false;
                    }
                }
                if (ok) {
                    const int r = r0 + v;
                    const int lp0 = __ldg(b.lane_ptr + 2 * r), lp1 = __ldg(b.lane_ptr + 2 * r + 1), lp2 = __ldg(b.lane_ptr + 2 * r + 2);
                    cols += (unsigned long long)(lp2 - lp0);
                    if (lanelet_side_sat<kWarp>(sm.bhx[warp], sm.bhy[warp], nbs, b.lane_x + lp0, b.lane_y + lp0, lp1 - lp0, t)) ok = false;
                    else if (lanelet_side_sat<kWarp>(sm.bhx[warp], sm.bhy[warp], nbs, b.lane_x + lp1, b.lane_y + lp1, lp2 - lp1, t)) ok = false;
                }
                if (ok && lane == 0) atomicOr(&sm.sv[v], 1u << i);
                __syncwarp();
            }
            __syncthreads();
            PROF_MARK(3);   // phase A
            // ---- phase B: vehicle-vs-vehicle tests (:37-44), one thread per (pair, i_u, i_v) -----------
            {
                int n_tests = 0, pofs[kJPairs + 1];
                for (int v = 1; v < nV; ++v)
                    for (int u = 0; u < v; ++u) {
                        pofs[jpair(u, v)] = n_tests;
                        n_tests += sm.nsucc[u] * sm.nsucc[v];
                    }
                for (int q = tid; q < n_tests; q += kJThreads) {
                    int u = 0, v = 1;
                    for (int vv = 1; vv < nV; ++vv)
                        for (int uu = 0; uu < vv; ++uu)
                            if (q >= pofs[jpair(uu, vv)]) { u = uu; v = vv; }
                    const int rel = q - pofs[jpair(u, v)];
                    const int iu = rel % sm.nsucc[u], iv = rel / sm.nsucc[u];
                    if (((sm.sv[u] >> iu) & 1u) && ((sm.sv[v] >> iv) & 1u) &&
                        sat_collide_scalar(sm.shx[u][iu], sm.shy[u][iu], sm.shn[u][iu], sm.shx[v][iv], sm.shy[v][iv], sm.shn[v][iv]))
                        atomicOr(&sm.pairbad[jpair(u, v)][iu], 1u << iv);
                }
            }
            __syncthreads();
            PROF_MARK(4);   // phase B
            // ---- phase C: the children, one thread each, pushed in child order ---------------------------
            for (long long c0 = 0; c0 < total; c0 += kJThreads) {
                const long long ci = c0 + tid;
                bool valid = false;
                double ef = 0.0;
                unsigned nid = 0;
                if (ci < total) {
                    nid = (unsigned)(n_nodes + 1 + ci);
                    long long rem = ci;
                    int iv[kMaxJoint];
                    double eg = cn.g, eh = 0.0;
                    valid = true;
#pragma unroll
                    for (int v = 0; v < kMaxJoint; ++v) {
                        if (v < nV) {
                            iv[v] = (int)(rem % sm.nsucc[v]);   // cartprod: first vehicle fastest
                            rem /= sm.nsucc[v];
                            eg = eg + sm.gterm[v][iv[v]];                                   // :61
                            for (int s2 = 0; s2 < to_go; ++s2) eh = eh + sm.hterm[v][iv[v]][s2];   // :66-73
                            if (!((sm.sv[v] >> iv[v]) & 1u)) valid = false;
                            for (int u = 0; u < v; ++u)
                                if ((sm.pairbad[jpair(u, v)][iv[u]] >> iv[v]) & 1u) valid = false;
                        }
                    }
                    JNode2 en;
                    en.g = eg; en.h = eh; en.pxs = xs; en.k = k_exp;
                    en.ci = (unsigned)ci | (valid ? 0u : 0x80000000u); en.pad = 0;
                    nn[nid] = en;                                                           // Tree.m:54-70 add_nodes
                    ef = eg + eh;                                                           // GraphSearch.m:102
                }
                const bool want = ci < total && (valid || exact);
                const unsigned bal = __ballot_sync(FULL, want);
                if (lane == 0) sm.wcount[warp] = __popc(bal);
                __syncthreads();
                const int nc0 = sm.ncand;
                int base = nc0, tot_round = 0;
                for (int w = 0; w < kJWarps; ++w) {
                    if (w < warp) base += sm.wcount[w];
                    tot_round += sm.wcount[w];
                }
                if (want) {
                    const int pos = base + __popc(bal & ((1u << lane) - 1u));
                    sm.cf[pos] = ef;
                    sm.cw[pos] = jpack(nid, (unsigned)k_exp) | (valid ? 0ULL : (1ULL << 63));
                }
                __syncthreads();
                const int ncand = nc0 + tot_round;
                const bool flush = ncand + kJThreads > kJCand || c0 + kJThreads >= total;
                if (flush) {
                    if (warp == 0) {   // :104 one pq.push per child, in order
                        for (int q0 = 0; q0 < ncand; q0 += kWarp) {
                            const int cnt = min(kWarp, ncand - q0);
                            HEnt he;
                            he.f = 0.0; he.w = 0;
                            if (lane < cnt) { he.f = sm.cf[q0 + lane]; he.w = sm.cw[q0 + lane]; }
                            heap.push_many(he, cnt, lane);
                        }
                    }
                    if (tid == 0) sm.ncand = 0;
                } else if (tid == 0) sm.ncand = ncand;
                __syncthreads();
            }
            n_nodes += (int)total;
            PROF_MARK(5);   // phase C + pushes
        }

        // ---- pops of the reference's queue (valid-only discipline): + the invalid nodes it met on the way
        const bool account = !exact && !tie && status == PDMPC_OK;
        if (account) {
            int extra = 0;
            bool amb = false;
            for (int i = 2 + tid; i <= n_nodes; i += kJThreads) {
                const JNode2 e = nn[i];
                if (e.ci & 0x80000000u) {
                    if (exhausted) ++extra;
                    else {
                        const double fi = e.g + e.h;
                        if (fi < f_goal) ++extra;
                        else if (fi == f_goal) amb = true;
                    }
                }
            }
            for (int d = 16; d > 0; d >>= 1) extra += __shfl_xor_sync(FULL, extra, d);
            amb = __any_sync(FULL, amb);
            if (lane == 0) sm.red[warp] = extra | (amb ? 0x40000000 : 0);
            __syncthreads();
            int ex_all = 0;
            for (int w = 0; w < kJWarps; ++w) {
                ex_all += sm.red[w] & 0x3fffffff;
                if (sm.red[w] & 0x40000000) tie = true;
            }
            __syncthreads();
            // (n_pops is kept by warp 0 only)
            n_pops += ex_all;
        }
        if (tie) {   // undecidable without the reference's tie mechanics: same search again, exact queue
            if (tid == 0) atomicAdd(o.counters + 3, 1ULL);
            redo_exact = true;
            __syncthreads();
            continue;
        }
        PROF_MARK(6);   // accounting
        PROF_FLUSH(o, tid == 0);
        if (status != PDMPC_OK) exhausted = true;
        // ---- results, row-wise (every vehicle of the search carries the shared fields) -------------------
        if (tid == 0) {
            unsigned cur = goal_xs;
            for (int d = Hp; d >= 0; --d) {
                sm.path_xs[d] = exhausted ? 0u : cur;
                if (!exhausted && d > 0) cur = xhdr(cur)->pxs;
            }
            atomicAdd(o.counters + 0, (unsigned long long)n_pops);
            atomicAdd(o.counters + 1, (unsigned long long)n_nodes);
        }
        if (lane == 0 && cols) atomicAdd(o.counters + 2, cols);   // (columns of the per-item tests: every lane counted them)
        __syncthreads();
        const double qnan = nan("");
        const int n_pops0 = __shfl_sync(FULL, n_pops, 0);
        const unsigned long long hash0 = __shfl_sync(FULL, hash, 0);
        if (warp == 0) {
            for (int v = 0; v < nV; ++v) {
                const int r = r0 + v;
                if (lane == 0) {
                    o.status[r] = status;
                    if (o.is_exhausted) o.is_exhausted[r] = exhausted ? 1 : 0;
                    if (o.n_expanded) o.n_expanded[r] = n_nodes;
                    if (o.n_pops) o.n_pops[r] = n_pops0;
                    if (o.pop_hash) o.pop_hash[r] = hash0;
                }
                for (int d = lane; d <= Hp; d += kWarp) {
                    const unsigned pxs = sm.path_xs[d];
                    const size_t oo = (size_t)r * (Hp + 1) + d;
                    JVeh pv;
                    pv.x = pv.y = pv.yaw = qnan; pv.trim = 0; pv.edge = 0;
                    double pg = qnan, ph = qnan;
                    unsigned pid = 0;
                    if (!exhausted) {
                        pv = xveh(pxs)[v];
                        pid = xhdr(pxs)->id;
                        const JNode2 pn = nn[pid];
                        pg = pn.g; ph = pn.h;
                    }
                    if (o.trims) o.trims[oo] = exhausted ? (d == 0 ? __ldg(b.trim0 + r) : 0) : (int)pv.trim;
                    if (o.tree_path) o.tree_path[oo] = (int)pid;
                    if (o.g_path) o.g_path[oo] = pg;
                    if (o.h_path) o.h_path[oo] = ph;
                    if (d >= 1) {
                        const size_t os = (size_t)r * Hp + (d - 1);
                        if (o.y_predicted) {
                            o.y_predicted[os * 3 + 0] = pv.x;
                            o.y_predicted[os * 3 + 1] = pv.y;
                            o.y_predicted[os * 3 + 2] = pv.yaw;
                        }
                        int ns = 0, edge = 0;
                        JVeh qv;
                        qv.x = qv.y = qv.c = qv.s = 0.0;
                        if (!exhausted) {
                            qv = xveh(sm.path_xs[d - 1])[v];
                            edge = (int)pv.edge;
                            ns = tb.area_npts[edge * 3 + PDMPC_AREA_NORMAL];
                        }
                        if (o.shape_npts) o.shape_npts[os] = ns;
                        if (o.shape_x && o.shape_y)
                            for (int i = 0; i < kAreaStride; ++i) {
                                double ox = 0.0, oy = 0.0;
                                if (i < ns) place_point(tb, edge, PDMPC_AREA_NORMAL, i, qv.c, qv.s, qv.x, qv.y, ox, oy);
                                o.shape_x[os * kAreaStride + i] = ox;
                                o.shape_y[os * kAreaStride + i] = oy;
                            }
                    }
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace pdmpc
