// pdmpc_cta.cuh — "one CTA per search" launch shape of the graph search: the lowest
// single-search latency, used when a batch has fewer searches than the GPU has SMs (one
// computation level of one 20-vehicle time step).
//
// Same algorithm and bit-identical results as search_kernel (pdmpc_kernels.cuh):
//   GraphSearch.do_graph_search / eval_edge_exact   hlc/optimizer/graph_search/GraphSearch.m:23-196
//   expand_node                                      hlc/optimizer/graph_search/expand_node.m:1-91
//   priority queue                                   .../priority_queue/priority_queue_interface_mex.cpp:19-108
//
// The reference validates an edge lazily, when its end node is popped (GraphSearch.m:64-77).
// Whether an edge is valid is a pure function of the node (parent pose, maneuver, depth), so
// the answer can be computed EARLY without changing any result:
//   * warp 0 ("master") owns the priority queue and the tree: pop -> read the node's validity
//     flag -> expand -> push.  Pop order, node ids and n_expanded are exactly the reference's:
//     invalid nodes are still pushed, popped and counted.
//   * warps 1..NH ("checkers") validate the children of every expansion as soon as they are
//     created, one child per warp (placement of the maneuver areas, InterX / SAT against the
//     staged obstacle polylines), compute cos/sin of a valid child's yaw for its own later
//     expansion, and publish a per-node flag in shared memory.
// The master therefore never runs an edge check; checks of nodes that are never popped are
// wasted work on SMs that would otherwise idle.
//
// VALID-ONLY QUEUE WITH DEFERRED INSERTION (`fast` launch argument, launch shape 5).  With the answers
// known early, an invalid child need not enter the queue at all, provided no pop ever has to break a tie:
//   * the reference's next VALID pop is a minimum of the whole queue, hence a minimum of the valid
//     entries; if that minimum is unique among the valid entries it is the same node whatever
//     invalid entries surround it, so the sequence of valid pops — and with it every expansion,
//     node id, n_expanded, the goal and the path — equals the reference's;
//   * new children wait in a PENDING buffer (one entry per lane of the master warp) until their flag
//     arrives: valid ones are then pushed, invalid ones dropped.  The master only blocks on a pending
//     child whose cost is not above the queue's current minimum — only such a child could be the next pop
//     (6.5 % of the pops of road-network records are children of the expansion before them); every other
//     check overlaps the master's pops.  (Round 1 waited for all children of every expansion: equal on
//     long searches, 1.6x slower on median ones, profiles/r01g_valid_only_queue.txt.)
//   * before every pop the minimum must be strictly below both children of the root.  On the first
//     non-unique minimum the search is RE-RUN from scratch with the exact queue (never observed on
//     road-network records: 0 of 4 M pops; 14 of 80 searches of the symmetric circle scenario);
//   * n_pops is recovered exactly: an invalid node was popped by the reference iff its f is below
//     the goal's (an equal f re-runs the search); all nodes are popped when the search exhausts.
//     pop_hash covers the valid pops only in this shape (documented in include/pdmpc_b200.h).
// A search then costs one heap pop per EXPANSION instead of one per created node cheaper than the goal
// (66 % of the pops of the longest road-network search are invalid nodes), on a queue a third the size.
//
// Hand-over: a ring of kRing job descriptors in shared memory; job j is published with
// bar.arrive on named barrier 1 + j % kRing, the checkers wait for it in bar.sync (no
// issue slots are spent spinning).  Node records the master needs again when a child is
// popped are kept in a direct-mapped shared-memory cache next to the HBM arena.
#pragma once

#include "pdmpc_kernels.cuh"

// Checker-side cycle accounting (build with -DPDMPC_PROFILE -DPDMPC_PROFILE_CHECKER: the counters
// of the master's phases then carry checker 0's per-child phases instead)
#if defined(PDMPC_PROFILE) && defined(PDMPC_PROFILE_CHECKER)
#define CPROF_DECL long long cprof_t0 = clock64(), cprof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define CPROF_MARK(i) do { long long _t = clock64(); cprof_acc[i] += _t - cprof_t0; cprof_t0 = _t; } while (0)
#define CPROF_FLUSH(o, on) do { if (on) for (int _i = 0; _i < 8; ++_i) atomicAdd((o).counters + 8 + _i, (unsigned long long)cprof_acc[_i]); } while (0)
#undef PROF_FLUSH
#define PROF_FLUSH(o, lane0)
#else
#define CPROF_DECL
#define CPROF_MARK(i)
#define CPROF_FLUSH(o, on)
#endif

namespace pdmpc {

constexpr int kRing = 8;            // job descriptors in flight (named barriers 1..kRing)
constexpr int kCtaHelpers = 12;     // checker warps (max branching of the single/triple-speed MPAs)
constexpr int kCtaHeap = 4096;      // heap entries in shared memory
constexpr int kCtaPts = 512;        // staged polyline points
constexpr int kCtaCache = 1024;     // node-record cache entries (direct mapped by id)
constexpr int kCtaFlags = 32768;    // validity flags (1 byte per node id) — searches need cap <= this
constexpr int kCtaDepCols = kDepCols;   // columns of predecessors' areas (pdmpc_plan_timestep)
constexpr int kCtaDepPolys = kCtaDepCols / kAreaStride;

struct __align__(16) CtaJob {       // one expansion: children nid0 .. nid0 + nchild - 1
    double px, py, pyaw, c, s;      // pose and cos/sin(yaw) of the expanded node
    unsigned nid0;
    int nchild, sbase, k;           // successor list base, depth of the children
    int terminate;
    int rot;                        // checker that takes child 0; child ci goes to checker (rot + ci) % NH
};

template <int HS, int SP, int NH, bool DEPS = false>
struct __align__(16) CtaSmem {
    double hf[HS + 2];               // heap costs, entry i at hf[i + 1] (pdmpc_heap_split.cuh)
    unsigned long long hw[HS];       // heap payloads
    double pts_x[SP], pts_y[SP];
    double refx[kMaxHp], refy[kMaxHp], vref[kMaxHp];
    double dmax[kMaxHp * kMaxHp];    // [k' * kMaxHp + t] = sum_{tau <= t} dt * v_ref(k' + tau), summed in order (expand_node.m:68)
    NodeA c_a[kCtaCache];
    NodeCS c_cs[kCtaCache];
    unsigned c_tag[kCtaCache];      // id whose (x, y, yaw, g) sits in c_a
    unsigned cs_tag[kCtaCache];     // id whose (cos, sin) sits in c_cs
    double shx[NH][kAreaStride], shy[NH][kAreaStride], bhx[NH][kAreaStride], bhy[NH][kAreaStride];
    CtaJob ring[kRing];
    int rng[kMaxHp + 2];
    unsigned path[kMaxHp + 1];
    unsigned done[NH];
    int abort_flag;
    unsigned next_search;
    int clear_upto;                 // highest node id of the previous search (flags to clear)
    int redo_exact;                 // the valid-only queue met a tie: run the same search again, exact queue
    int mode_exact;
    unsigned char flag[kCtaFlags];  // 0 pending, 1 valid, 2 invalid
    // pdmpc_plan_timestep: the areas the predecessors of this search have planned, kAreaStride
    // columns each (points, then NaN: the separator column of vectorize_all_obstacles.m:36-63 and
    // padding that no InterX inequality can satisfy); predecessor r at step k sits at
    // ((k - 1) * n_pred + r) * kAreaStride
    double dep_x[DEPS ? kCtaDepCols : 1], dep_y[DEPS ? kCtaDepCols : 1];
    int dep_n[DEPS ? kCtaDepPolys : 1];   // points of each of them (SAT needs the exact count)
    int dep_cols[kMaxHp + 1];             // real columns per step (statistics only)
    int n_pred;
};

// Ordering between the master and the checkers (all flags, descriptors and cache entries live in
// shared memory; the only global data handed over is cos/sin of a node, same SM).  An acquire-
// release fence at CTA scope is enough; __threadfence_block() compiles to the sequentially
// consistent MEMBAR.SC.CTA, which drains the master's outstanding arena stores on every pop.
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Warps of a CTA are spread round-robin over the SM's four schedulers, so warps 4, 8, 12, ...
// share the master's.  They are left idle (parked at the CTA barrier): the master's dependent
// instruction chain then never waits for an issue slot behind a checker's FP64 stream
// (profiles/r01e_cta_latency.txt).
template <int HS, int SP, int NH, bool DEPS = false>
__global__ void __launch_bounds__((NH + NH / 3) * kWarp, 1)
search_cta_kernel(MpaDev m, BatchDev b, OutDev o, ArenaDev ar, unsigned *work_counter, int heap_smem, int fast,
                  DepsDev dp) {
    // ids whose checks can be in flight: kRing jobs x at most PDMPC_MAX_TRIMS - 1 children each, so the
    // checkers of two nodes that share a cache slot (ids kCtaCache apart) never run at the same time
    static_assert(kRing * PDMPC_MAX_TRIMS <= kCtaCache, "a late checker must never alias a newer cache entry");
    constexpr int TILE = kWarp;
    static_assert(NH % 3 == 0, "three checker warps per scheduler group");
    constexpr int kThreads = (NH + NH / 3) * kWarp;        // launched: master + checkers + parked warps
    constexpr int kBarThreads = (NH + 1) * kWarp;          // master + checkers: the named barriers' count
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CtaSmem<HS, SP, NH, DEPS> &sm = *reinterpret_cast<CtaSmem<HS, SP, NH, DEPS> *>(smem_raw);
    Tables tb;
    tb.succ_ptr = m.succ_ptr; tb.succ_te = m.succ_te; tb.edge_d = m.edge_d;
    tb.area_npts = m.area_npts; tb.area_x = m.area_x; tb.area_y = m.area_y;

    const int warp_id = threadIdx.x / kWarp;
    Tile<TILE> t;
    t.shift = 0; t.lane = threadIdx.x % kWarp; t.mask = 0xffffffffu;
    const int Hp = m.Hp, nT = m.nT;
    const size_t slot_base = (size_t)blockIdx.x * (size_t)ar.cap;
    NodeA *__restrict__ na = ar.a + slot_base;
    NodeB *__restrict__ nb = ar.b + slot_base;
    NodeCS *__restrict__ ncs = ar.cs + slot_base;
    volatile unsigned char *vflag = sm.flag;
    volatile unsigned *vdone = sm.done;
    volatile int *vabort = &sm.abort_flag;
    volatile unsigned *vcs_tag = sm.cs_tag;
    if (threadIdx.x == 0) { sm.clear_upto = kCtaFlags - 1; sm.redo_exact = 0; }   // all flags, the first time
    const unsigned hf_addr = shared_base_once(sm.hf), hw_addr = shared_base_once(sm.hw);

    for (;;) {
        // ---- fetch a search; cooperative set-up ------------------------------------------
        __syncthreads();
        if (threadIdx.x == 0) {
            if (sm.redo_exact) { sm.redo_exact = 0; sm.mode_exact = 1; }   // same search again
            else { sm.next_search = atomicAdd(work_counter, 1u); sm.mode_exact = fast ? 0 : 1; }
        }
        __syncthreads();
        const unsigned si_u = sm.next_search;
        const bool exact = sm.mode_exact != 0;
        if (si_u >= (unsigned)b.n) break;
        const int si = b.order ? __ldg(b.order + si_u) : (int)si_u;
        const int clear_upto = sm.clear_upto;
        for (int i = threadIdx.x; i <= clear_upto / 4; i += kThreads) reinterpret_cast<unsigned *>(sm.flag)[i] = 0u;
        for (int i = threadIdx.x; i < kCtaCache; i += kThreads) { sm.c_tag[i] = 0u; sm.cs_tag[i] = 0u; }
        if (threadIdx.x < NH) sm.done[threadIdx.x] = 0u;
        if (threadIdx.x == 0) sm.abort_flag = 0;
        for (int k = threadIdx.x; k < Hp; k += kThreads) {
            sm.refx[k] = __ldg(b.ref_x + (size_t)si * Hp + k);
            sm.refy[k] = __ldg(b.ref_y + (size_t)si * Hp + k);
            sm.vref[k] = __ldg(b.v_ref + (size_t)si * Hp + k);
        }
        if (threadIdx.x >= 1 && threadIdx.x < Hp) {   // d_traveled_max of expand_node.m:68 for children of step k'
            const int kx = threadIdx.x;
            double d = 0.0;
            for (int it = 1; it <= Hp - kx; ++it) {
                d = d + b.dt * __ldg(b.v_ref + (size_t)si * Hp + kx + it - 1);
                sm.dmax[kx * kMaxHp + it] = d;
            }
        }
        const int *slot = b.slot_ptr + (size_t)si * (Hp + 1);
        const int trim0 = __ldg(b.trim0 + si);
        const int sp0 = __ldg(slot + 0), sp1 = __ldg(slot + 1);
        const int lp0 = __ldg(b.lane_ptr + 2 * si), lp1 = __ldg(b.lane_ptr + 2 * si + 1),
                  lp2 = __ldg(b.lane_ptr + 2 * si + 2);
        const double *opx = nullptr, *opy = nullptr, *lpx = nullptr, *lpy = nullptr;
        int obase = 0, llo = 0, lhi = 0;
        if (b.checker == PDMPC_CHECKER_INTERX) {
            const int spE = __ldg(slot + Hp + 1);
            const int ob_lo = __ldg(b.poly_ptr + sp0) + sp0, ob_hi = __ldg(b.poly_ptr + spE) + spE;
            const int ll_lo = lp0 + 2 * si, ll_hi = lp2 + 2 * si + 2;
            for (int k = threadIdx.x; k <= Hp + 1; k += kThreads) {
                const int q = __ldg(slot + k);
                sm.rng[k] = __ldg(b.poly_ptr + q) + q - ob_lo;
            }
            const int nl = ll_hi - ll_lo, no = ob_hi - ob_lo;
            int used = 0;
            if (nl <= SP) {
                for (int j = threadIdx.x; j < nl; j += kThreads) {
                    sm.pts_x[j] = __ldg(b.ll_x + ll_lo + j);
                    sm.pts_y[j] = __ldg(b.ll_y + ll_lo + j);
                }
                lpx = sm.pts_x; lpy = sm.pts_y; llo = 0; lhi = nl;
                used = nl;
            } else {
                lpx = b.ll_x; lpy = b.ll_y; llo = ll_lo; lhi = ll_hi;
            }
            if (used + no <= SP) {
                for (int j = threadIdx.x; j < no; j += kThreads) {
                    sm.pts_x[used + j] = __ldg(b.pl_x + ob_lo + j);
                    sm.pts_y[used + j] = __ldg(b.pl_y + ob_lo + j);
                }
                opx = sm.pts_x; opy = sm.pts_y; obase = used;
            } else {
                opx = b.pl_x; opy = b.pl_y; obase = ob_lo;
            }
        }
        // (everything above is independent of the predecessors: it overlaps their searches)
        if (DEPS) {
            // ---- consider_predecessors (PrioritizedController.m:449-506) on the device ------------
            // Work items are handed out in a topological order of the DAG (the host sorts them), so
            // every predecessor's ticket is below this one's: it is finished or running on another
            // CTA, never waiting behind this CTA.
            const int q0 = __ldg(dp.pred_ptr + si), q1 = __ldg(dp.pred_ptr + si + 1);
            const int npred = q1 - q0;
            for (int r = threadIdx.x; r < npred; r += kThreads) {
                const int *flag = dp.done + __ldg(dp.pred_idx + q0 + r);
                while (ld_acquire_gpu(flag) == 0) __nanosleep(100);
            }
            __syncthreads();
            if (threadIdx.x == 0) sm.n_pred = npred;
            if (threadIdx.x <= Hp) sm.dep_cols[threadIdx.x] = 0;
            __syncthreads();
            const double qn = nan("");
            for (int idx = threadIdx.x; idx < npred * Hp * kAreaStride; idx += kThreads) {
                const int v = idx % kAreaStride, pk = idx / kAreaStride;   // pk = (k - 1) * npred + r
                const int r = pk % npred, k0 = pk / npred;
                const int j = __ldg(dp.pred_idx + q0 + r);
                const bool planned = ld_acquire_gpu(dp.done + j) == 1;
                const size_t os = (size_t)j * Hp + k0;
                int np = 0;
                double vx = qn, vy = qn;
                if (planned) {   // what j has just written: read from L2, never through the read-only path
                    np = __ldcg(o.shape_npts + os);
                    if (v < np) { vx = __ldcg(o.shape_x + os * kAreaStride + v); vy = __ldcg(o.shape_y + os * kAreaStride + v); }
                } else if (dp.fb_npts) {
                    np = __ldg(dp.fb_npts + os);
                    if (v < np) { vx = __ldg(dp.fb_x + os * kAreaStride + v); vy = __ldg(dp.fb_y + os * kAreaStride + v); }
                }
                sm.dep_x[idx] = vx; sm.dep_y[idx] = vy;
                if (v == 0) {
                    sm.dep_n[pk] = np;
                    if (np) atomicAdd(&sm.dep_cols[k0 + 1], np + 1);
                }
            }
        }
        __syncthreads();

        if (warp_id != 0 && (warp_id & 3) == 0) continue;   // parked warp: back to the CTA barrier
        if (warp_id != 0) {
            // =================== checker warps: eval_edge_exact, eagerly ===================
            const int w = warp_id - 1 - (warp_id >> 2);
            double *shx = sm.shx[w], *shy = sm.shy[w], *bhx = sm.bhx[w], *bhy = sm.bhy[w];
            unsigned long long cols = 0;
            CPROF_DECL
            for (unsigned j = 0;; ++j) {
                CPROF_MARK(7);
                named_bar_sync(1 + (int)(j % kRing), kBarThreads);          // job j is published
                CPROF_MARK(0);   // waiting for a job
#ifdef PDMPC_PROFILE
                const long long ck0 = clock64();
#endif
                const CtaJob &jb = sm.ring[j % kRing];
                if (jb.terminate) break;
                const int nchild = jb.nchild, cK = jb.k;
                if (!*vabort) {
                    const double ppx = jb.px, ppy = jb.py, pc = jb.c, ps = jb.s, pyaw = jb.pyaw;
                    const unsigned nid0 = jb.nid0;
                    const int sbase = jb.sbase;
                    // consecutive expansions start at different checkers: with 3-4 children per expansion and
                    // up to kRing jobs in flight, a fixed child -> checker map would serialise every job on
                    // checkers 0..3 (measured: checker 0 busy for the whole search, the master waiting on it)
                    for (int ci = (w - jb.rot + NH) % NH; ci < nchild; ci += NH) {
                        CPROF_MARK(7);
                        const int te = tb.succ_te[sbase + ci];
                        const int edge = te >> 8;
                        const unsigned nid = nid0 + (unsigned)ci;
                        const int bkind = (cK == Hp) ? PDMPC_AREA_LARGE_OFFSET : PDMPC_AREA_WITHOUT_OFFSET;  // GraphSearch.m:166-174
                        const int ns = tb.area_npts[edge * 3 + PDMPC_AREA_NORMAL];
                        const int nbs = tb.area_npts[edge * 3 + bkind];
                        if (t.lane < 8)
                            place_point(tb, edge, PDMPC_AREA_NORMAL, t.lane, pc, ps, ppx, ppy, shx[t.lane], shy[t.lane]);
                        else if (t.lane < 16)
                            place_point(tb, edge, bkind, t.lane - 8, pc, ps, ppx, ppy, bhx[t.lane - 8], bhy[t.lane - 8]);
                        t.sync();
                        CPROF_MARK(1);   // table loads + placement
                        bool valid = true;
                        if (b.checker == PDMPC_CHECKER_INTERX) {
                            const int st_lo = obase + sm.rng[0], st_hi = obase + sm.rng[1];
                            const int dy_lo = obase + sm.rng[cK], dy_hi = obase + sm.rng[cK + 1];
                            cols += (unsigned long long)((st_hi - st_lo) + (dy_hi - dy_lo) + (lhi - llo));
                            if (DEPS) cols += (unsigned long long)sm.dep_cols[cK];
                            const bool hit_obs = interx_dispatch<TILE>(ns, opx, opy, st_lo, st_hi, dy_lo, dy_hi, shx, shy, t);
                            CPROF_MARK(2);   // InterX against the obstacles
                            if (hit_obs)
                                valid = false;
                            else if (DEPS && sm.n_pred > 0 &&
                                     interx_dispatch<TILE>(ns, sm.dep_x, sm.dep_y, (cK - 1) * sm.n_pred * kAreaStride,
                                                           cK * sm.n_pred * kAreaStride, 0, 0, shx, shy, t))
                                valid = false;
                            else if (interx_dispatch<TILE>(nbs, lpx, lpy, llo, lhi, 0, 0, bhx, bhy, t))
                                valid = false;
                        } else {
                            const int dp0 = __ldg(slot + cK), dp1 = __ldg(slot + cK + 1);
                            for (int pass = 0; pass < 2 && valid; ++pass) {
                                const int q0 = pass == 0 ? sp0 : dp0, q1 = pass == 0 ? sp1 : dp1;
                                for (int p = q0; p < q1 && valid; ++p) {
                                    const int v0 = __ldg(b.poly_ptr + p), v1 = __ldg(b.poly_ptr + p + 1);
                                    cols += (unsigned long long)(v1 - v0);
                                    if (sat_collide<TILE>(shx, shy, ns, b.vert_x + v0, b.vert_y + v0, v1 - v0, t))
                                        valid = false;
                                }
                            }
                            if (DEPS) {
                                const int np_ = sm.n_pred;
                                for (int r = 0; r < np_ && valid; ++r) {
                                    const int pk = (cK - 1) * np_ + r, nv = sm.dep_n[pk];
                                    if (nv < 2) continue;
                                    cols += (unsigned long long)nv;
                                    if (sat_collide<TILE, false>(shx, shy, ns, sm.dep_x + pk * kAreaStride,
                                                                 sm.dep_y + pk * kAreaStride, nv, t))
                                        valid = false;
                                }
                            }
                            if (valid) {
                                cols += (unsigned long long)(lp2 - lp0);
                                if (lanelet_side_sat<TILE>(bhx, bhy, nbs, b.lane_x + lp0, b.lane_y + lp0, lp1 - lp0, t))
                                    valid = false;
                                else if (lanelet_side_sat<TILE>(bhx, bhy, nbs, b.lane_x + lp1, b.lane_y + lp1, lp2 - lp1, t))
                                    valid = false;
                            }
                        }
                        CPROF_MARK(3);   // InterX against predecessors' areas + lanelet bounds (or SAT)
                        if (t.lane == 0) {
                            if (valid && cK < Hp) {
                                // cos/sin of the child's yaw for ITS expansion (expand_node.m:50-51);
                                // yaw' = yaw + dyaw exactly as the master computes it (:55)
                                NodeCS ecs;
                                sincos_ref(pyaw + tb.edge_d[edge * 4 + 2], ecs.s, ecs.c);
                                ncs[nid] = ecs;
                                if (exact) {
                                    // cache entry, seqlock style: the master may be reading this slot for
                                    // an older node (id - kCtaCache) right now.  (The valid-only queue knows
                                    // the answer before the pop and loads the arena record early instead.)
                                    const int cslot = nid & (kCtaCache - 1);
                                    vcs_tag[cslot] = 0u;
                                    fence_cta();
                                    sm.c_cs[cslot] = ecs;
                                    fence_cta();
                                    vcs_tag[cslot] = nid;
                                }
                            }
                            fence_cta();
                            vflag[nid] = valid ? 1 : 2;
                        }
                        CPROF_MARK(4);   // sincos + publication
                        t.sync();   // shapes are rewritten by the next child
                    }
                }
                t.sync();   // every lane is done with the descriptor
#ifdef PDMPC_PROFILE
                if (t.lane == 0 && w == 0) { atomicAdd(o.counters + 4, (unsigned long long)(clock64() - ck0)); atomicAdd(o.counters + 5, 1ULL); }
#endif
                if (t.lane == 0) vdone[w] = j + 1u;
            }
            if (t.lane == 0 && cols) atomicAdd(o.counters + 2, cols);
            CPROF_FLUSH(o, t.lane == 0 && w == 0);
            continue;   // next search (meets the master at the __syncthreads on top)
        }

        // =================== master warp: queue + tree ==========================================
        HeapSplit heap;
        heap.sf = hf_addr; heap.sw = hw_addr;
        heap.gl = ar.heap + slot_base;
        heap.hs = heap_smem;       // <= HS; smaller values only exercise the arena overflow (tests)
        heap.len = 1;
        int n_nodes = 1, n_pops = 0, status = PDMPC_OK;
        unsigned long long hash = 0xcbf29ce484222325ULL;
        bool exhausted = false, tie = false;
        unsigned goal = 0, n_jobs = 0;
        int rot = 0;
        double f_last = 0.0;
        // valid-only queue: children whose flag is not known yet, one per lane
        bool pocc = false;
        double pf = 0.0;
        unsigned long long pw = 0;
        // flags of the pending children: valid -> pushed, invalid -> dropped (cost parked in the unused cos/sin
        // record for the n_pops accounting).  Blocks while `need(pf)` holds for a child without an answer.
        auto resolve_pending = [&](bool wait_all) {
            for (;;) {
                const unsigned pid_ = (unsigned)(pw & 0x1fffffu);
                const unsigned fl = pocc ? (unsigned)vflag[pid_] : 0u;
                fence_cta();
                if (pocc && fl == 2u) { NodeCS park; park.c = pf; park.s = 0.0; ncs[pid_] = park; pocc = false; }
                const bool okv = pocc && fl == 1u;
                const unsigned vm = __ballot_sync(0xffffffffu, okv);
                if (vm) {
                    const int mv = __popc(vm);
                    unsigned mm = vm;
                    for (int i = 0; i < t.lane && i < mv; ++i) mm &= mm - 1u;
                    const int src = t.lane < mv ? __ffs(mm) - 1 : 0;
                    HEnt hv;
                    hv.f = __shfl_sync(0xffffffffu, pf, src);
                    hv.w = __shfl_sync(0xffffffffu, pw, src);
                    heap.push_many(hv, mv, t.lane);
                    if (okv) pocc = false;
                }
                // only a pending child whose cost is not above the queue's minimum can be the next pop
                const bool some = heap.len > 0;
                const double thr = some ? heap.f_at(0) : 0.0;
                const bool must = pocc && (wait_all || !some || !(pf > thr));
                if (!__any_sync(0xffffffffu, must)) break;
                __nanosleep(20);
            }
        };
        if (t.lane == 0) {   // root: GraphSearch.m:34-46
            NodeA ra;
            ra.x = __ldg(b.x0 + si); ra.y = __ldg(b.y0 + si); ra.yaw = __ldg(b.yaw0 + si); ra.g = 0.0;
            NodeCS rcs;
            sincos_ref(ra.yaw, rcs.s, rcs.c);
            NodeB rb;
            rb.h = 0.0; rb.parent = 0; rb.edge = 0xffff; rb.trim = (unsigned char)trim0; rb.k = 0;
            na[1] = ra; nb[1] = rb; ncs[1] = rcs;
            sm.c_a[1] = ra; sm.c_tag[1] = 1u; sm.c_cs[1] = rcs; sm.cs_tag[1] = 1u;
            HEnt re;
            re.f = 0.0; re.w = HEnt::pack(1u, 0u, 0u, 0u, (unsigned)trim0);
            heap.st(0, re);
        }
        t.sync();

        PROF_DECL
        for (;;) {   // GraphSearch.m:53-107
            PROF_MARK(7);
            if (!exact) {
                resolve_pending(false);
                PROF_MARK(6);   // pending children: flags, pushes, waits
            }
            if (heap.len == 0) { exhausted = true; break; }               // :57-61 (no pending child is left either)
            if (!exact && !heap.min_is_unique()) { tie = true; break; }   // tie mechanics would matter: re-run exact
            // valid-only queue: the node about to be popped is valid, its records are complete: fetch them now,
            // their latency hides behind the heap walk
            NodeA ca_e = {0.0, 0.0, 0.0, 0.0};
            NodeCS ccs_e = {0.0, 0.0};
            int sb_e = 0, nc_e = 0;
            if (!exact) {
                const unsigned long long w0 = lds_u64(heap.sw);
                const unsigned id0 = (unsigned)(w0 & 0x1fffffu);
                const int k0 = (int)((w0 >> 52) & 0x1fu), tr0 = (int)(w0 >> 57) + 1;
                // (L2 loads: the checkers' cos/sin stores come from other warps of this SM, never rely on L1)
                const double2 a0 = __ldcg(reinterpret_cast<const double2 *>(na + id0));
                const double2 a1 = __ldcg(reinterpret_cast<const double2 *>(na + id0) + 1);
                const double2 c0_ = __ldcg(reinterpret_cast<const double2 *>(ncs + id0));
                ca_e.x = a0.x; ca_e.y = a0.y; ca_e.yaw = a1.x; ca_e.g = a1.y;
                ccs_e.c = c0_.x; ccs_e.s = c0_.y;
                if (k0 < Hp) {
                    sb_e = tb.succ_ptr[k0 * nT + (tr0 - 1)];
                    nc_e = tb.succ_ptr[k0 * nT + (tr0 - 1) + 1] - sb_e;
                }
            }
            const HEnt top = heap.pop<true>(t.lane);
            PROF_MARK(1);   // heap pop
            const unsigned id = top.id(), par = top.pid();
            const int cK = (int)top.k();
            ++n_pops;
            if (!fast) hash = hash_step(hash, id);   // shape 4: every pop; shape 5: valid pops only (below)
            f_last = top.f;
            if (exact && par != 0) {   // eval_edge_exact's answer, computed by a checker warp
                unsigned f;
                do { f = vflag[id]; } while (f == 0u);
                fence_cta();
                PROF_MARK(2);   // wait for the checker's answer
                if (f != 1u) continue;                                    // :75-77
            }
            if (fast) hash = hash_step(hash, id);
            if (cK == Hp) { goal = id; break; }                           // :81-90

            // ---- expand_node.m:1-91 (nV == 1) ----------------------------------------------
            const int ctrim = (int)top.trim();
            const int k_exp = cK + 1;
            const int sbase = exact ? tb.succ_ptr[(k_exp - 1) * nT + (ctrim - 1)] : sb_e;
            const int nchild = exact ? tb.succ_ptr[(k_exp - 1) * nT + (ctrim - 1) + 1] - sbase : nc_e;
            if (n_nodes + nchild >= ar.cap || n_nodes + nchild >= kCtaFlags) { status = PDMPC_ERR_CAPACITY; break; }
            const int cslot = id & (kCtaCache - 1);
            NodeA ca = ca_e;
            NodeCS ccs = ccs_e;
            if (exact) {
                if (sm.c_tag[cslot] == id) ca = sm.c_a[cslot]; else ca = na[id];
                const unsigned t1 = vcs_tag[cslot];
                fence_cta();
                const volatile double *vcs = reinterpret_cast<const volatile double *>(&sm.c_cs[cslot]);
                ccs.c = vcs[0]; ccs.s = vcs[1];
                fence_cta();
                const unsigned t2 = vcs_tag[cslot];
                if (t1 != id || t2 != id) ccs = ncs[id];   // evicted (or being replaced): HBM arena copy
            }
            const double s = ccs.s, c = ccs.c;
            // publish the job first: the checkers work while the master computes costs and pushes
            {
                while (true) {   // ring slot free: every checker is done with job n_jobs - kRing
                    const unsigned d = t.lane < NH ? vdone[t.lane] : n_jobs;   // jobs completed by checker `lane`
                    if (__all_sync(0xffffffffu, d + kRing > n_jobs)) break;
                }
                if (t.lane == 0) {
                    CtaJob &jb = sm.ring[n_jobs % kRing];
                    jb.px = ca.x; jb.py = ca.y; jb.pyaw = ca.yaw; jb.c = c; jb.s = s;
                    jb.nid0 = (unsigned)(n_nodes + 1); jb.nchild = nchild; jb.sbase = sbase; jb.k = k_exp;
                    jb.terminate = 0; jb.rot = rot;
                }
                rot = (rot + nchild) % NH;
                fence_cta();
                t.sync();
                named_bar_arrive(1 + (int)(n_jobs % kRing), kBarThreads);
                ++n_jobs;
            }
            PROF_MARK(3);   // node record + job hand-over
            const int to_go = Hp - k_exp;               // :37
            for (int c0 = 0; c0 < nchild; c0 += TILE) {
                const int ci = c0 + t.lane;
                const int cnt = min(TILE, nchild - c0);
                HEnt he;
                he.f = 0.0; he.w = 0;
                const unsigned nid = (unsigned)(n_nodes + 1 + ci);
                if (ci < nchild) {
                    const int te = tb.succ_te[sbase + ci];
                    const int cedge = te >> 8, t2 = (te & 0xff) + 1;
                    const double mdx = tb.edge_d[cedge * 4 + 0], mdy = tb.edge_d[cedge * 4 + 1],
                                 mdyaw = tb.edge_d[cedge * 4 + 2];
                    NodeA ea;
                    ea.x = c * mdx - s * mdy + ca.x;          // :53
                    ea.y = s * mdx + c * mdy + ca.y;          // :54
                    ea.yaw = ca.yaw + mdyaw;                  // :55
                    const double ddx = ea.x - sm.refx[k_exp - 1], ddy = ea.y - sm.refy[k_exp - 1];
                    const double nrm = sqrt(ddx * ddx + ddy * ddy);
                    ea.g = ca.g + nrm * nrm;            // :61
                    double eh = 0.0;                    // :66-73
                    for (int it0 = 1; it0 <= to_go; it0 += 6) {
                        double hn[6];                   // independent square roots, issued together
#pragma unroll
                        for (int u = 0; u < 6; ++u) {
                            const int kk = min(k_exp + it0 + u - 1, Hp - 1);
                            const double hx = ea.x - sm.refx[kk], hy = ea.y - sm.refy[kk];
                            hn[u] = sqrt(hx * hx + hy * hy);
                        }
#pragma unroll
                        for (int u = 0; u < 6; ++u) {
                            if (it0 + u <= to_go) {
                                const double mm = fmax(0.0, hn[u] - sm.dmax[k_exp * kMaxHp + it0 + u]);
                                eh = eh + mm * mm;
                            }
                        }
                    }
                    NodeB eb;
                    eb.h = eh; eb.parent = id; eb.edge = (unsigned short)cedge;
                    eb.trim = (unsigned char)t2; eb.k = (unsigned char)k_exp;
                    na[nid] = ea;                       // Tree.m:54-70 add_nodes
                    nb[nid] = eb;
                    const int nslot = nid & (kCtaCache - 1);
                    sm.c_a[nslot] = ea;
                    sm.c_tag[nslot] = nid;
                    he.f = ea.g + eh;                   // GraphSearch.m:102 (weights 1)
                    he.w = HEnt::pack(nid, id, (unsigned)cedge, (unsigned)k_exp, (unsigned)t2);
                }
                PROF_MARK(4);   // successor generation
                if (exact) {
                    heap.push_many(he, cnt, t.lane);    // :104, one push per child, in order
                } else {
                    // valid-only queue: the children wait in the pending buffer for their flags
                    while (__popc(__ballot_sync(0xffffffffu, !pocc)) < cnt) resolve_pending(true);
                    const unsigned freem = __ballot_sync(0xffffffffu, !pocc);
                    const int r = __popc(freem & ((1u << t.lane) - 1u));   // this lane's rank among the free lanes
                    const double nf = __shfl_sync(0xffffffffu, he.f, r & 31);
                    const unsigned long long nw = __shfl_sync(0xffffffffu, he.w, r & 31);
                    if (!pocc && r < cnt) { pocc = true; pf = nf; pw = nw; }
                }
                PROF_MARK(5);   // heap pushes
            }
            n_nodes += nchild;
        }
        PROF_FLUSH(o, t.lane == 0);

        // ---- release the checkers, then write the results (GraphSearch.m:58-60 / :82-89) ------
        if (!exact && !tie && status == PDMPC_OK) {
            // the n_pops accounting below needs the flag of every node cheaper than the goal: let the checkers
            // finish the published jobs, then take the answers of the children still pending
            while (true) {
                const unsigned d = t.lane < NH ? vdone[t.lane] : n_jobs;
                if (__all_sync(0xffffffffu, d >= n_jobs)) break;
                __nanosleep(32);
            }
            fence_cta();
            resolve_pending(true);
        }
        if (t.lane == 0) *vabort = 1;
        {
            while (true) {
                const unsigned d = t.lane < NH ? vdone[t.lane] : n_jobs;   // jobs completed by checker `lane`
                if (__all_sync(0xffffffffu, d + kRing > n_jobs)) break;
            }
            if (t.lane == 0) sm.ring[n_jobs % kRing].terminate = 1;
            fence_cta();
            t.sync();
            named_bar_arrive(1 + (int)(n_jobs % kRing), kBarThreads);
        }
        if (status != PDMPC_OK) exhausted = true;
        if (!exact && !tie && status == PDMPC_OK) {
            // pops of the reference's queue = valid pops + the invalid nodes it met on the way:
            // all of them if the search exhausted, those cheaper than the goal otherwise
            int extra = 0;
            bool amb = false;
            for (int i0 = 2; i0 <= n_nodes; i0 += TILE) {
                const int i = i0 + t.lane;
                if (i <= n_nodes && vflag[i] == 2u) {
                    if (exhausted) ++extra;
                    else {
                        const double fi = ncs[i].c;
                        if (fi < f_last) ++extra;
                        else if (fi == f_last) amb = true;
                    }
                }
            }
            for (int d = 16; d > 0; d >>= 1) extra += __shfl_xor_sync(0xffffffffu, extra, d);
            if (__any_sync(0xffffffffu, amb)) tie = true;
            n_pops += extra;
        }
        if (tie) {   // undecidable without the reference's tie mechanics: same search again, exact queue
            if (t.lane == 0) { sm.redo_exact = 1; sm.clear_upto = n_nodes; atomicAdd(o.counters + 3, 1ULL); }
            continue;
        }
        if (t.lane == 0) {
            unsigned cur = goal;
            for (int d = Hp; d >= 0; --d) {           // Tree.m:44-52 path_to_root, flipped
                sm.path[d] = exhausted ? 0u : cur;
                if (!exhausted && d > 0) cur = nb[cur].parent;
            }
            o.status[si] = status;
            if (o.is_exhausted) o.is_exhausted[si] = exhausted ? 1 : 0;
            if (o.n_expanded) o.n_expanded[si] = n_nodes;
            if (o.n_pops) o.n_pops[si] = n_pops;
            if (o.pop_hash) o.pop_hash[si] = hash;
            atomicAdd(o.counters + 0, (unsigned long long)n_pops);
            atomicAdd(o.counters + 1, (unsigned long long)n_nodes);
        }
        t.sync();
        const double qnan = nan("");
        for (int d = t.lane; d <= Hp; d += TILE) {
            const unsigned pid = sm.path[d];
            NodeA pa = {qnan, qnan, qnan, qnan};
            NodeB pb;
            pb.h = qnan; pb.parent = 0; pb.edge = 0; pb.trim = 0; pb.k = 0;
            if (!exhausted) { pa = na[pid]; pb = nb[pid]; }
            const size_t oo = (size_t)si * (Hp + 1) + d;
            if (o.trims) o.trims[oo] = exhausted ? (d == 0 ? trim0 : 0) : (int)pb.trim;
            if (o.tree_path) o.tree_path[oo] = (int)pid;
            if (o.g_path) o.g_path[oo] = pa.g;
            if (o.h_path) o.h_path[oo] = pb.h;
            if (d >= 1) {
                const size_t os = (size_t)si * Hp + (d - 1);
                if (o.y_predicted) {                   // return_path_to.m:11-25
                    o.y_predicted[os * 3 + 0] = pa.x;
                    o.y_predicted[os * 3 + 1] = pa.y;
                    o.y_predicted[os * 3 + 2] = pa.yaw;
                }
                if (o.shape_npts) {                    // return_path_area.m:5-7
                    int edge = 0, ns = 0;
                    NodeA qa = {0.0, 0.0, 0.0, 0.0};
                    NodeCS qcs = {0.0, 0.0};
                    if (!exhausted) {
                        const unsigned qid = sm.path[d - 1];   // parent on the path (valid, depth < Hp)
                        qa = na[qid];
                        qcs = ncs[qid];
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
            }
        }
        if (DEPS) {   // publish_predictions (PrioritizedController.m:355-364): the areas above are final
            __threadfence();
            t.sync();
            if (t.lane == 0) st_release_gpu(dp.done + si, exhausted ? 2 : 1);
        }
        if (t.lane == 0) sm.clear_upto = n_nodes;
    }
}

}  // namespace pdmpc
