/*
 * pdmpc_oracle.c — CPU oracle (TEST INFRASTRUCTURE, see pdmpc_oracle.h).
 *
 * Plain-C, IEEE-double restatement of the reference's per-vehicle graph search.
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).  Every
 * function cites the reference file:line it follows (paths relative to the
 * reference checkout).
 */
#include "pdmpc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* sin / cos: the reference calls MATLAB's cos/sin (GraphSearch.m:155-156,
 * expand_node.m:50-51), closed source.  Oracle and device path both use the
 * algorithm below so that costs, and therefore pop order,
 * agree bit for bit.  |error| < 1.5 ulp for |x| < 2^20 * pi/2.               */
static const double SC_TWO_OVER_PI = 0x1.45f306dc9c883p-1;
static const double SC_P1 = 0x1.921fb54400000p+0;  /* pi/2, leading 33 bits */
static const double SC_P2 = 0x1.0b4611a600000p-34; /* next 33 bits          */
static const double SC_P3 = 0x1.3198a2e037073p-69; /* remainder             */
static const double SC_S[10] = {
    -0x1.5555555555555p-3, 0x1.1111111111111p-7, -0x1.a01a01a01a01ap-13,
    0x1.71de3a556c734p-19, -0x1.ae64567f544e4p-26, 0x1.6124613a86d09p-33,
    -0x1.ae7f3e733b81fp-41, 0x1.952c77030ad4ap-49, -0x1.2f49b46814157p-57,
    0x1.71b8ef6dcf572p-66};
static const double SC_C[10] = {
    -0x1.0000000000000p-1, 0x1.5555555555555p-5, -0x1.6c16c16c16c17p-10,
    0x1.a01a01a01a01ap-16, -0x1.27e4fb7789f5cp-22, 0x1.1eed8eff8d898p-29,
    -0x1.93974a8c07c9dp-37, 0x1.ae7f3e733b81fp-45, -0x1.6827863b97d97p-53,
    0x1.e542ba4020225p-62};

void oracle_sincos(double x, double *s, double *c) {
    double n = rint(x * SC_TWO_OVER_PI);
    double r = ((x - n * SC_P1) - n * SC_P2) - n * SC_P3;
    double z = r * r;
    double ps = SC_S[9];
    double pc = SC_C[9];
    for (int i = 8; i >= 0; --i) {
        ps = ps * z + SC_S[i];
        pc = pc * z + SC_C[i];
    }
    double sr = r + (r * z) * ps; /* sin(r) */
    double cr = 1.0 + z * pc;     /* cos(r) */
    long long q = (long long)n & 3LL;
    switch (q) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
    }
}

/* ------------------------------------------------------------------------- */
/* intersect_sat.m:17-42 (intersect_a_b)                                      */
static int sat_a_b(const double *x1, const double *y1, int n1,
                   const double *x2, const double *y2, int n2) {
    /* :19 edge_vector = diff([shape1, shape1(:,1)],1,2) -> n1 edges incl. closing */
    for (int e = 0; e < n1; ++e) {
        int e1 = (e + 1 == n1) ? 0 : e + 1;
        double ex = x1[e1] - x1[e];
        double ey = y1[e1] - y1[e];
        /* :21 axis = [-ey; ex]; :23 normed = axis ./ vecnorm(axis) */
        double ax = -ey, ay = ex;
        double nrm = sqrt(ax * ax + ay * ay);
        double nx = ax / nrm, ny = ay / nrm; /* zero edge -> NaN axis */
        /* :26-32 projections; MATLAB min/max skip NaN == C fmin/fmax */
        double mn1 = NAN, mx1 = NAN, mn2 = NAN, mx2 = NAN;
        for (int v = 0; v < n1; ++v) {
            double d = nx * x1[v] + ny * y1[v];
            mn1 = fmin(mn1, d);
            mx1 = fmax(mx1, d);
        }
        for (int v = 0; v < n2; ++v) {
            double d = nx * x2[v] + ny * y2[v];
            mn2 = fmin(mn2, d);
            mx2 = fmax(mx2, d);
        }
        /* :33-40 */
        double d1 = mn1 - mx2;
        double d2 = mn2 - mx1;
        if (d1 > 0 || d2 > 0) return 0;
    }
    return 1;
}

/* intersect_sat.m:1-15 */
int oracle_intersect_sat(const double *x1, const double *y1, int n1,
                         const double *x2, const double *y2, int n2) {
    if (!sat_a_b(x1, y1, n1, x2, y2, n2)) return 0;
    if (!sat_a_b(x2, y2, n2, x1, y1, n1)) return 0;
    return 1;
}

/* intersect_lanelets.m:1-22: per segment right bound first, then left bound */
int oracle_intersect_lanelets(const double *sx, const double *sy, int ns,
                              const double *rx, const double *ry,
                              const double *lx, const double *ly, int n) {
    for (int i = 0; i + 1 < n; ++i) {
        if (oracle_intersect_sat(sx, sy, ns, rx + i, ry + i, 2)) return 1;
        if (oracle_intersect_sat(sx, sy, ns, lx + i, ly + i, 2)) return 1;
    }
    return 0;
}

/* intersect_lanelet_boundary.m:1-56 */
static int lanelet_side(const double *sx, const double *sy, int ns, double max_x,
                        double min_x, double max_y, double min_y,
                        const double *bx, const double *by, int nb) {
    for (int n = 0; n + 1 < nb; ++n) {
        double ax = bx[n], bx2 = bx[n + 1], ay = by[n], by2 = by[n + 1];
        /* :20 / :40 all(max_x < seg_x) || all(min_x > seg_x) || ... */
        if ((max_x < ax && max_x < bx2) || (min_x > ax && min_x > bx2) ||
            (max_y < ay && max_y < by2) || (min_y > ay && min_y > by2))
            continue;
        if (oracle_intersect_sat(sx, sy, ns, bx + n, by + n, 2)) return 1;
    }
    return 0;
}

int oracle_intersect_lanelet_boundary(const double *sx, const double *sy, int ns,
                                      const double *lx, const double *ly, int nl,
                                      const double *rx, const double *ry, int nr) {
    /* :11-14 max/min over shape (NaN-free input) */
    double max_x = sx[0], min_x = sx[0], max_y = sy[0], min_y = sy[0];
    for (int i = 1; i < ns; ++i) {
        max_x = fmax(max_x, sx[i]);
        min_x = fmin(min_x, sx[i]);
        max_y = fmax(max_y, sy[i]);
        min_y = fmin(min_y, sy[i]);
    }
    if (lanelet_side(sx, sy, ns, max_x, min_x, max_y, min_y, lx, ly, nl)) return 1;
    if (lanelet_side(sx, sy, ns, max_x, min_x, max_y, min_y, rx, ry, nr)) return 1;
    return 0;
}

/* InterX.m:48-85,108-110, isReturnPoints == false */
int oracle_interx(const double *x1, const double *y1, int n1,
                  const double *x2, const double *y2, int n2) {
    if (n1 == 0 || n2 == 0) return 0; /* :48-52 */
    for (int i = 0; i + 1 < n1; ++i) {
        double dx1 = x1[i + 1] - x1[i]; /* :63 */
        double dy1 = y1[i + 1] - y1[i];
        double S1 = dx1 * y1[i] - dy1 * x1[i]; /* :67 */
        for (int j = 0; j + 1 < n2; ++j) {
            double dx2 = x2[j + 1] - x2[j]; /* :64 */
            double dy2 = y2[j + 1] - y2[j];
            double S2 = dx2 * y2[j] - dy2 * x2[j]; /* :68 */
            /* :70  C1 = D(dx1*y2 - dy1*x2, S1) < 0 */
            double a0 = (dx1 * y2[j] - dy1 * x2[j]) - S1;
            double a1 = (dx1 * y2[j + 1] - dy1 * x2[j + 1]) - S1;
            int c1 = (a0 * a1) < 0;
            /* :71  C2 = (D((y1*dx2 - x1*dy2)', S2') < 0)' */
            double b0 = (y1[i] * dx2 - x1[i] * dy2) - S2;
            double b1 = (y1[i + 1] * dx2 - x1[i + 1] * dy2) - S2;
            int c2 = (b0 * b1) < 0;
            if (c1 && c2) return 1; /* :74-85 */
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* priority queue: priority_queue_interface_mex.cpp:19-31 (min on value) on
 * top of libstdc++ stl_heap.h (__push_heap :135-147, __adjust_heap :224-249). */
typedef struct {
    int64_t id;
    double val;
} pq_entry;
struct oracle_pq {
    pq_entry *a;
    int64_t len, cap;
};

/* comp(a,b) of the reference: value(a) > value(b) */
static inline int pq_comp(const pq_entry *a, const pq_entry *b) { return a->val > b->val; }

static void pq_push_heap(pq_entry *first, int64_t hole, int64_t top, pq_entry value) {
    int64_t parent = (hole - 1) / 2;
    while (hole > top && pq_comp(first + parent, &value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

static void pq_adjust_heap(pq_entry *first, int64_t hole, int64_t len, pq_entry value) {
    const int64_t top = hole;
    int64_t second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (pq_comp(first + second, first + (second - 1))) second--;
        first[hole] = first[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        first[hole] = first[second - 1];
        hole = second - 1;
    }
    pq_push_heap(first, hole, top, value);
}

oracle_pq *oracle_pq_new(void) {
    oracle_pq *q = (oracle_pq *)calloc(1, sizeof(*q));
    return q;
}
void oracle_pq_free(oracle_pq *q) {
    if (!q) return;
    free(q->a);
    free(q);
}
void oracle_pq_push(oracle_pq *q, int64_t id, double value) {
    if (q->len == q->cap) {
        q->cap = q->cap ? q->cap * 2 : 256;
        q->a = (pq_entry *)realloc(q->a, (size_t)q->cap * sizeof(pq_entry));
    }
    pq_entry e = {id, value};
    q->a[q->len++] = e;                       /* c.push_back(x) */
    pq_push_heap(q->a, q->len - 1, 0, e);     /* std::push_heap */
}
/* diagnostic: pops whose minimum was not unique (a child of the root had the same value) */
static int64_t g_tie_pops = 0;
int64_t oracle_tie_pops(int reset) {
    int64_t v = __atomic_load_n(&g_tie_pops, __ATOMIC_RELAXED);
    if (reset) __atomic_store_n(&g_tie_pops, 0, __ATOMIC_RELAXED);
    return v;
}

int64_t oracle_pq_pop(oracle_pq *q, double *value) {
    if (q->len == 0) return -1;               /* ...mex.cpp:87-94 */
    pq_entry top = q->a[0];
    if ((q->len > 1 && q->a[1].val == top.val) || (q->len > 2 && q->a[2].val == top.val))
        __atomic_fetch_add(&g_tie_pops, 1, __ATOMIC_RELAXED);
    if (q->len > 1) {                         /* std::pop_heap */
        pq_entry last = q->a[q->len - 1];
        q->a[q->len - 1] = q->a[0];
        pq_adjust_heap(q->a, 0, q->len - 1, last);
    }
    q->len--;                                 /* c.pop_back() */
    if (value) *value = top.val;
    return top.id;
}
int64_t oracle_pq_size(const oracle_pq *q) { return q->len; }

/* ------------------------------------------------------------------------- */
/* Tree.m: SoA arena, ids 1-based (index 0 unused so ids index directly).     */
typedef struct {
    double *x, *y, *yaw, *g, *h;
    int32_t *trim, *k, *parent;
    int64_t size, cap; /* size == number of nodes == tree.size() */
} tree_t;

static void tree_reserve(tree_t *t, int64_t need) {
    if (need + 1 <= t->cap) return;
    int64_t cap = t->cap ? t->cap : 1024;
    while (cap < need + 1) cap *= 2;
    t->x = (double *)realloc(t->x, (size_t)cap * sizeof(double));
    t->y = (double *)realloc(t->y, (size_t)cap * sizeof(double));
    t->yaw = (double *)realloc(t->yaw, (size_t)cap * sizeof(double));
    t->g = (double *)realloc(t->g, (size_t)cap * sizeof(double));
    t->h = (double *)realloc(t->h, (size_t)cap * sizeof(double));
    t->trim = (int32_t *)realloc(t->trim, (size_t)cap * sizeof(int32_t));
    t->k = (int32_t *)realloc(t->k, (size_t)cap * sizeof(int32_t));
    t->parent = (int32_t *)realloc(t->parent, (size_t)cap * sizeof(int32_t));
    t->cap = cap;
}
static void tree_free(tree_t *t) {
    free(t->x); free(t->y); free(t->yaw); free(t->g); free(t->h);
    free(t->trim); free(t->k); free(t->parent);
    memset(t, 0, sizeof(*t));
}

typedef struct {
    tree_t tree;
    oracle_pq pq;
    double *ox, *oy; /* vectorize_all_obstacles scratch */
    int64_t ocap;
    int32_t *edge_of; /* n_trims*n_trims -> edge index or -1 */
} work_t;

/* FNV-1a over the popped ids taken as 32-bit words (parity trace of the pop order) */
static uint64_t fnv1a_u32(uint64_t h, uint32_t v) { return (h ^ (uint64_t)v) * 0x100000001b3ULL; }

/* rotate/translate a maneuver area: GraphSearch.m:158-159 (and :162-163,:168-169) */
static void place_area(const pdmpc_mpa_desc *mpa, int edge, int kind, double c, double s,
                       double px, double py, double *ox, double *oy, int *n) {
    int np = mpa->area_npts[edge * 3 + kind];
    const double *ax = mpa->area_x + (size_t)(edge * 3 + kind) * PDMPC_AREA_STRIDE;
    const double *ay = mpa->area_y + (size_t)(edge * 3 + kind) * PDMPC_AREA_STRIDE;
    for (int i = 0; i < np; ++i) {
        ox[i] = c * ax[i] - s * ay[i] + px;
        oy[i] = s * ax[i] + c * ay[i] + py;
    }
    *n = np;
}

/* are_constraints_satisfied_sat.m:1-68, nV == 1.  The HDV block (:55-66) is
 * unreachable in the reference (`if ~any(adj)` then loop over find(adj)). */
static int constraints_sat(const pdmpc_batch_in *in, int Hp, int si, int k,
                           const double *sx, const double *sy, int ns,
                           const double *bx, const double *by, int nb) {
    const int32_t *slot = in->slot_ptr + (size_t)si * (Hp + 1);
    /* :15-22 static obstacles, then :24-35 dynamic obstacles of step k */
    for (int pass = 0; pass < 2; ++pass) {
        int s = pass == 0 ? 0 : k;
        for (int p = slot[s]; p < slot[s + 1]; ++p) {
            int v0 = in->poly_ptr[p], v1 = in->poly_ptr[p + 1];
            if (oracle_intersect_sat(sx, sy, ns, in->vert_x + v0, in->vert_y + v0, v1 - v0))
                return 0;
        }
    }
    /* :46-53 lanelet boundary with the boundary-check shape.  Empty boundary
     * cells make both loops of intersect_lanelet_boundary run zero times. */
    int l0 = in->lane_ptr[2 * si], l1 = in->lane_ptr[2 * si + 1], l2 = in->lane_ptr[2 * si + 2];
    if (oracle_intersect_lanelet_boundary(bx, by, nb, in->lane_x + l0, in->lane_y + l0, l1 - l0,
                                          in->lane_x + l1, in->lane_y + l1, l2 - l1))
        return 0;
    return 1;
}

/* vectorize_all_obstacles.m:36-63: [static..., dynamic(:,k)...], each polygon
 * followed by a [NaN;NaN] column.  Returns the column count. */
static int64_t vectorize_step(const pdmpc_batch_in *in, int Hp, int si, int k, work_t *w) {
    const int32_t *slot = in->slot_ptr + (size_t)si * (Hp + 1);
    int64_t need = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int s = pass == 0 ? 0 : k;
        for (int p = slot[s]; p < slot[s + 1]; ++p) need += in->poly_ptr[p + 1] - in->poly_ptr[p] + 1;
    }
    if (need > w->ocap) {
        w->ocap = need * 2 + 64;
        w->ox = (double *)realloc(w->ox, (size_t)w->ocap * sizeof(double));
        w->oy = (double *)realloc(w->oy, (size_t)w->ocap * sizeof(double));
    }
    int64_t n = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int s = pass == 0 ? 0 : k;
        for (int p = slot[s]; p < slot[s + 1]; ++p) {
            for (int v = in->poly_ptr[p]; v < in->poly_ptr[p + 1]; ++v) {
                w->ox[n] = in->vert_x[v];
                w->oy[n] = in->vert_y[v];
                ++n;
            }
            w->ox[n] = NAN;
            w->oy[n] = NAN;
            ++n;
        }
    }
    return n;
}

/* are_constraints_satisfied_interx.m:1-39, no HDVs (hdv_obstacles{k} empty ->
 * is_hdv_obstacle false). */
static int constraints_interx(const pdmpc_batch_in *in, int Hp, int si, int k, work_t *w,
                              const double *sx, const double *sy, int ns,
                              const double *bx, const double *by, int nb,
                              const double *lanex, const double *laney, int nlane) {
    int64_t nobs = vectorize_step(in, Hp, si, k, w);
    if (oracle_interx(sx, sy, ns, w->ox, w->oy, (int)nobs)) return 0; /* :17 */
    if (oracle_interx(bx, by, nb, lanex, laney, nlane)) return 0;      /* :34 */
    return 1;
}

typedef struct {
    int64_t *trace;
    int64_t cap, n;
} trace_t;

/* Parity-trace option: hash only the popped nodes that pass their edge check (what launch
 * shape 5 of the CUDA path reports, include/pdmpc_b200.h).  Default: every popped node. */
static int g_hash_valid_only = 0;
void oracle_set_hash_valid_pops_only(int on) { g_hash_valid_only = on; }

/* GraphSearch.m:23-109 for search `si`. */
static int search_one(const pdmpc_mpa_desc *mpa, const pdmpc_batch_in *in, pdmpc_batch_out *out,
                      int si, work_t *w, trace_t *tr) {
    const int Hp = mpa->Hp, nT = mpa->n_trims;
    tree_t *t = &w->tree;
    oracle_pq *pq = &w->pq;
    pq->len = 0;
    t->size = 0;
    tree_reserve(t, 1);
    /* :34-41 root */
    t->size = 1;
    t->x[1] = in->x0[si];
    t->y[1] = in->y0[si];
    t->yaw[1] = in->yaw0[si];
    t->trim[1] = in->trim0[si];
    t->k[1] = 0;
    t->g[1] = 0;
    t->h[1] = 0;
    t->parent[1] = 0;
    oracle_pq_push(pq, 1, 0.0); /* :45-46 */

    /* :48-49 set_up_constraints: lanelet = [left, NaN, right, NaN]
     * (vectorize_all_obstacles.m:27-30); built for both checkers, used by InterX */
    int l0 = in->lane_ptr[2 * si], l1 = in->lane_ptr[2 * si + 1], l2 = in->lane_ptr[2 * si + 2];
    int nlane = (l2 - l0) + 2;
    double *lanex = (double *)malloc((size_t)nlane * sizeof(double));
    double *laney = (double *)malloc((size_t)nlane * sizeof(double));
    {
        int n = 0;
        for (int i = l0; i < l1; ++i) { lanex[n] = in->lane_x[i]; laney[n] = in->lane_y[i]; ++n; }
        lanex[n] = NAN; laney[n] = NAN; ++n;
        for (int i = l1; i < l2; ++i) { lanex[n] = in->lane_x[i]; laney[n] = in->lane_y[i]; ++n; }
        lanex[n] = NAN; laney[n] = NAN; ++n;
    }

    const double *rx = in->ref_x + (size_t)si * Hp, *ry = in->ref_y + (size_t)si * Hp;
    const double *vr = in->v_ref + (size_t)si * Hp;
    int64_t n_pops = 0;
    uint64_t hash = 0xcbf29ce484222325ULL;
    int exhausted = 0;
    int64_t goal = 0;

    for (;;) {
        int64_t id = oracle_pq_pop(pq, NULL); /* :55 */
        if (id == -1) {                       /* :57-61 */
            exhausted = 1;
            break;
        }
        ++n_pops;
        if (!g_hash_valid_only) hash = fnv1a_u32(hash, (uint32_t)id);
        if (tr && tr->n < tr->cap) tr->trace[tr->n++] = id;

        /* :64-73 eval_edge_exact (:111-196) */
        int is_valid = 1;
        int32_t par = t->parent[id];
        if (par) { /* :137-139 root is valid without any check */
            double pX = t->x[par], pY = t->y[par], pYaw = t->yaw[par];
            int t1 = t->trim[par], t2 = t->trim[id], cK = t->k[id];
            int edge = w->edge_of[(t1 - 1) * nT + (t2 - 1)];
            double c, s;
            oracle_sincos(pYaw, &s, &c); /* :155-156 */
            double sx[PDMPC_AREA_STRIDE], sy[PDMPC_AREA_STRIDE];
            double bx[PDMPC_AREA_STRIDE], by[PDMPC_AREA_STRIDE];
            int ns, nb;
            place_area(mpa, edge, PDMPC_AREA_NORMAL, c, s, pX, pY, sx, sy, &ns); /* :158-160 */
            /* :166-174 boundary shape: large offset at k == Hp, else without offset */
            place_area(mpa, edge, cK == Hp ? PDMPC_AREA_LARGE_OFFSET : PDMPC_AREA_WITHOUT_OFFSET,
                       c, s, pX, pY, bx, by, &nb);
            if (in->checker == PDMPC_CHECKER_SAT)
                is_valid = constraints_sat(in, Hp, si, cK, sx, sy, ns, bx, by, nb);
            else
                is_valid = constraints_interx(in, Hp, si, cK, w, sx, sy, ns, bx, by, nb, lanex,
                                              laney, nlane);
        }
        if (!is_valid) continue; /* :75-77 */
        if (g_hash_valid_only) hash = fnv1a_u32(hash, (uint32_t)id);

        if (t->k[id] == Hp) { /* :81-90 */
            goal = id;
            break;
        }

        /* :93-104 expand_node.m:1-91, nV == 1 */
        {
            double curX = t->x[id], curY = t->y[id], curYaw = t->yaw[id], curG = t->g[id];
            int curTrim = t->trim[id];
            int k_exp = t->k[id] + 1; /* :13 */
            const uint8_t *row = mpa->transition + ((size_t)(k_exp - 1) * nT + (curTrim - 1)) * nT;
            int time_steps_to_go = Hp - k_exp; /* :37 */
            double c, s;
            oracle_sincos(curYaw, &s, &c); /* :50-51 */
            int64_t first_new = t->size + 1;
            for (int j = 0; j < nT; ++j) { /* :18 find(...) ascending */
                if (!row[j]) continue;
                int edge = w->edge_of[(curTrim - 1) * nT + j];
                double dx = mpa->edge_dx[edge], dy = mpa->edge_dy[edge], dyaw = mpa->edge_dyaw[edge];
                double ex = c * dx - s * dy + curX; /* :53 */
                double ey = s * dx + c * dy + curY; /* :54 */
                double eyaw = curYaw + dyaw;        /* :55 */
                /* :61 g + norm([..])^2, reference point k_exp */
                double ddx = ex - rx[k_exp - 1], ddy = ey - ry[k_exp - 1];
                double nrm = sqrt(ddx * ddx + ddy * ddy);
                double eg = curG + nrm * nrm;
                /* :66-73 */
                double eh = 0, d_traveled_max = 0;
                for (int it = 1; it <= time_steps_to_go; ++it) {
                    d_traveled_max = d_traveled_max + in->dt_seconds * vr[k_exp + it - 1];
                    double hx = ex - rx[k_exp + it - 1], hy = ey - ry[k_exp + it - 1];
                    double hn = sqrt(hx * hx + hy * hy);
                    double m = fmax(0.0, hn - d_traveled_max);
                    eh = eh + m * m;
                }
                /* Tree.m:54-70 add_nodes */
                tree_reserve(t, t->size + 1);
                int64_t nid = ++t->size;
                t->x[nid] = ex; t->y[nid] = ey; t->yaw[nid] = eyaw;
                t->trim[nid] = j + 1; t->k[nid] = k_exp;
                t->g[nid] = eg; t->h[nid] = eh; t->parent[nid] = (int32_t)id;
            }
            /* :100-104 push in order, value = g*1 + h*1 */
            for (int64_t nid = first_new; nid <= t->size; ++nid)
                oracle_pq_push(pq, nid, t->g[nid] * 1 + t->h[nid] * 1);
        }
    }
    free(lanex);
    free(laney);

    /* results: GraphSearch.m:58-60 / :82-89 */
    if (out->status) out->status[si] = PDMPC_OK;
    if (out->is_exhausted) out->is_exhausted[si] = (uint8_t)exhausted;
    if (out->n_expanded) out->n_expanded[si] = (int32_t)t->size;
    if (out->n_pops) out->n_pops[si] = (int32_t)n_pops;
    if (out->pop_hash) out->pop_hash[si] = hash;
    int64_t path[PDMPC_MAX_HP + 1];
    if (!exhausted) { /* Tree.m:44-52 path_to_root, flipped */
        int64_t n = goal;
        for (int d = Hp; d >= 0; --d) {
            path[d] = n;
            n = t->parent[n];
        }
    }
    for (int d = 0; d <= Hp; ++d) {
        size_t o = (size_t)si * (Hp + 1) + d;
        if (out->trims) out->trims[o] = exhausted ? (d == 0 ? in->trim0[si] : 0) : t->trim[path[d]];
        if (out->tree_path) out->tree_path[o] = exhausted ? 0 : (int32_t)path[d];
        if (out->g_path) out->g_path[o] = exhausted ? NAN : t->g[path[d]];
        if (out->h_path) out->h_path[o] = exhausted ? NAN : t->h[path[d]];
    }
    for (int d = 1; d <= Hp; ++d) {
        size_t o = (size_t)si * Hp + (d - 1);
        if (out->y_predicted) { /* return_path_to.m:11-25 */
            out->y_predicted[o * 3 + 0] = exhausted ? NAN : t->x[path[d]];
            out->y_predicted[o * 3 + 1] = exhausted ? NAN : t->y[path[d]];
            out->y_predicted[o * 3 + 2] = exhausted ? NAN : t->yaw[path[d]];
        }
        if (out->shape_npts) { /* return_path_area.m:5-7: shapes stored at pop (GraphSearch.m:79) */
            double sx[PDMPC_AREA_STRIDE], sy[PDMPC_AREA_STRIDE];
            int ns = 0;
            for (int i = 0; i < PDMPC_AREA_STRIDE; ++i) sx[i] = sy[i] = 0.0;
            if (!exhausted) {
                int64_t par = path[d - 1], ch = path[d];
                int edge = w->edge_of[(t->trim[par] - 1) * nT + (t->trim[ch] - 1)];
                double c, s;
                oracle_sincos(t->yaw[par], &s, &c);
                place_area(mpa, edge, PDMPC_AREA_NORMAL, c, s, t->x[par], t->y[par], sx, sy, &ns);
            }
            out->shape_npts[o] = ns;
            if (out->shape_x && out->shape_y)
                for (int i = 0; i < PDMPC_AREA_STRIDE; ++i) {
                    out->shape_x[o * PDMPC_AREA_STRIDE + i] = sx[i];
                    out->shape_y[o * PDMPC_AREA_STRIDE + i] = sy[i];
                }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Centralized (joint) search: the same GraphSearch.do_graph_search with iter.amount = nV > 1
 * (CentralizedController.m:33-59).  Rows of the batch: r = search * nV + vehicle; the obstacle slots of a
 * search are those of its vehicle-0 row.  expand_node.m:15-26 (Cartesian product of the vehicles'
 * successors, first vehicle fastest: cartprod), :43-75 (costs summed over the vehicles in order),
 * GraphSearch.m:150-192 (vehicles checked in order, first failure ends the edge check),
 * are_constraints_satisfied_sat.m:15-53 (static, dynamic, vehicles i < iVeh, own lanelet boundary). */
typedef struct {
    double *x, *y, *yaw, *g, *h;   /* x, y, yaw: [node * nV + v] */
    int32_t *trim, *k, *parent;    /* trim: [node * nV + v] */
    int64_t size, cap;
} jtree_t;

static void jtree_reserve(jtree_t *t, int64_t need, int nV) {
    if (need + 1 <= t->cap) return;
    int64_t cap = t->cap ? t->cap : 4096;
    while (cap < need + 1) cap *= 2;
    t->x = (double *)realloc(t->x, (size_t)cap * nV * sizeof(double));
    t->y = (double *)realloc(t->y, (size_t)cap * nV * sizeof(double));
    t->yaw = (double *)realloc(t->yaw, (size_t)cap * nV * sizeof(double));
    t->trim = (int32_t *)realloc(t->trim, (size_t)cap * nV * sizeof(int32_t));
    t->g = (double *)realloc(t->g, (size_t)cap * sizeof(double));
    t->h = (double *)realloc(t->h, (size_t)cap * sizeof(double));
    t->k = (int32_t *)realloc(t->k, (size_t)cap * sizeof(int32_t));
    t->parent = (int32_t *)realloc(t->parent, (size_t)cap * sizeof(int32_t));
    t->cap = cap;
}

static int joint_search_one(const pdmpc_mpa_desc *mpa, const pdmpc_batch_in *in, pdmpc_batch_out *out,
                            int s, int nV, const int32_t *edge_of, int64_t max_nodes) {
    const int Hp = mpa->Hp, nT = mpa->n_trims;
    const int r0 = s * nV;
    jtree_t tt;
    memset(&tt, 0, sizeof(tt));
    jtree_t *t = &tt;
    oracle_pq pq;
    memset(&pq, 0, sizeof(pq));
    jtree_reserve(t, 1, nV);
    t->size = 1;
    for (int v = 0; v < nV; ++v) {   /* GraphSearch.m:34-41 */
        t->x[1 * nV + v] = in->x0[r0 + v];
        t->y[1 * nV + v] = in->y0[r0 + v];
        t->yaw[1 * nV + v] = in->yaw0[r0 + v];
        t->trim[1 * nV + v] = in->trim0[r0 + v];
    }
    t->k[1] = 0; t->g[1] = 0; t->h[1] = 0; t->parent[1] = 0;
    oracle_pq_push(&pq, 1, 0.0);
    int64_t n_pops = 0;
    uint64_t hash = 0xcbf29ce484222325ULL;
    int exhausted = 0, status = PDMPC_OK;
    int64_t goal = 0;
    const int32_t *slot = in->slot_ptr + (size_t)r0 * (Hp + 1);
    double sx[PDMPC_MAX_JOINT][PDMPC_AREA_STRIDE], sy[PDMPC_MAX_JOINT][PDMPC_AREA_STRIDE];
    int nsv[PDMPC_MAX_JOINT];

    for (;;) {
        int64_t id = oracle_pq_pop(&pq, NULL);
        if (id == -1) { exhausted = 1; break; }
        ++n_pops;
        if (!g_hash_valid_only) hash = fnv1a_u32(hash, (uint32_t)id);
        int is_valid = 1;
        int32_t par = t->parent[id];
        if (par) {
            int cK = t->k[id];
            for (int v = 0; v < nV && is_valid; ++v) {   /* GraphSearch.m:150-192 */
                double pX = t->x[par * nV + v], pY = t->y[par * nV + v], pYaw = t->yaw[par * nV + v];
                int edge = edge_of[(t->trim[par * nV + v] - 1) * nT + (t->trim[id * nV + v] - 1)];
                double c, sn;
                oracle_sincos(pYaw, &sn, &c);
                double bx[PDMPC_AREA_STRIDE], by[PDMPC_AREA_STRIDE];
                int nb;
                place_area(mpa, edge, PDMPC_AREA_NORMAL, c, sn, pX, pY, sx[v], sy[v], &nsv[v]);
                place_area(mpa, edge, cK == Hp ? PDMPC_AREA_LARGE_OFFSET : PDMPC_AREA_WITHOUT_OFFSET, c, sn, pX, pY,
                           bx, by, &nb);
                /* are_constraints_satisfied_sat.m:15-35 */
                for (int pass = 0; pass < 2 && is_valid; ++pass) {
                    int q = pass == 0 ? 0 : cK;
                    for (int p = slot[q]; p < slot[q + 1]; ++p) {
                        int v0 = in->poly_ptr[p], v1 = in->poly_ptr[p + 1];
                        if (oracle_intersect_sat(sx[v], sy[v], nsv[v], in->vert_x + v0, in->vert_y + v0, v1 - v0)) {
                            is_valid = 0;
                            break;
                        }
                    }
                }
                /* :37-44 vehicles of the same node with a lower index */
                for (int u = v - 1; u >= 0 && is_valid; --u)
                    if (oracle_intersect_sat(sx[u], sy[u], nsv[u], sx[v], sy[v], nsv[v])) is_valid = 0;
                /* :46-53 own lanelet boundary */
                if (is_valid) {
                    int r = r0 + v;
                    int l0 = in->lane_ptr[2 * r], l1 = in->lane_ptr[2 * r + 1], l2 = in->lane_ptr[2 * r + 2];
                    if (oracle_intersect_lanelet_boundary(bx, by, nb, in->lane_x + l0, in->lane_y + l0, l1 - l0,
                                                          in->lane_x + l1, in->lane_y + l1, l2 - l1))
                        is_valid = 0;
                }
            }
        }
        if (!is_valid) continue;
        if (g_hash_valid_only) hash = fnv1a_u32(hash, (uint32_t)id);
        if (t->k[id] == Hp) { goal = id; break; }

        /* expand_node.m */
        int k_exp = t->k[id] + 1;
        int succ[PDMPC_MAX_JOINT][PDMPC_MAX_TRIMS], nsucc[PDMPC_MAX_JOINT];
        int64_t total = 1;
        for (int v = 0; v < nV; ++v) {   /* :17-26 */
            const uint8_t *row = mpa->transition + ((size_t)(k_exp - 1) * nT + (t->trim[id * nV + v] - 1)) * nT;
            nsucc[v] = 0;
            for (int j = 0; j < nT; ++j)
                if (row[j]) succ[v][nsucc[v]++] = j + 1;
            total *= nsucc[v];
        }
        if (t->size + total >= max_nodes) { status = PDMPC_ERR_CAPACITY; exhausted = 1; break; }
        int time_steps_to_go = Hp - k_exp;
        jtree_reserve(t, t->size + total, nV);
        double cs[PDMPC_MAX_JOINT], sn[PDMPC_MAX_JOINT];
        for (int v = 0; v < nV; ++v) oracle_sincos(t->yaw[id * nV + v], &sn[v], &cs[v]);   /* :50-51 */
        int64_t first_new = t->size + 1;
        for (int64_t ci = 0; ci < total; ++ci) {   /* cartprod: first vehicle fastest */
            int64_t nid = ++t->size, rem = ci;
            double eg = t->g[id], eh = 0;
            for (int v = 0; v < nV; ++v) {
                int t2 = succ[v][rem % nsucc[v]];
                rem /= nsucc[v];
                int r = r0 + v;
                const double *rx = in->ref_x + (size_t)r * Hp, *ry = in->ref_y + (size_t)r * Hp;
                const double *vr = in->v_ref + (size_t)r * Hp;
                int edge = edge_of[(t->trim[id * nV + v] - 1) * nT + (t2 - 1)];
                double dx = mpa->edge_dx[edge], dy = mpa->edge_dy[edge], dyaw = mpa->edge_dyaw[edge];
                double cX = t->x[id * nV + v], cY = t->y[id * nV + v], cYaw = t->yaw[id * nV + v];
                double ex = cs[v] * dx - sn[v] * dy + cX;
                double ey = sn[v] * dx + cs[v] * dy + cY;
                double ddx = ex - rx[k_exp - 1], ddy = ey - ry[k_exp - 1];
                double nrm = sqrt(ddx * ddx + ddy * ddy);
                eg = eg + nrm * nrm;   /* :61 */
                double d_traveled_max = 0;
                for (int it = 1; it <= time_steps_to_go; ++it) {   /* :66-73 */
                    d_traveled_max = d_traveled_max + in->dt_seconds * vr[k_exp + it - 1];
                    double hx = ex - rx[k_exp + it - 1], hy = ey - ry[k_exp + it - 1];
                    double m = fmax(0.0, sqrt(hx * hx + hy * hy) - d_traveled_max);
                    eh = eh + m * m;
                }
                t->x[nid * nV + v] = ex; t->y[nid * nV + v] = ey; t->yaw[nid * nV + v] = cYaw + dyaw;
                t->trim[nid * nV + v] = t2;
            }
            t->k[nid] = k_exp; t->g[nid] = eg; t->h[nid] = eh; t->parent[nid] = (int32_t)id;
        }
        for (int64_t nid = first_new; nid <= t->size; ++nid) oracle_pq_push(&pq, nid, t->g[nid] * 1 + t->h[nid] * 1);
    }

    int64_t path[PDMPC_MAX_HP + 1];
    if (!exhausted) {
        int64_t n = goal;
        for (int d = Hp; d >= 0; --d) { path[d] = n; n = t->parent[n]; }
    }
    for (int v = 0; v < nV; ++v) {
        int r = r0 + v;
        if (out->status) out->status[r] = status;
        if (out->is_exhausted) out->is_exhausted[r] = (uint8_t)exhausted;
        if (out->n_expanded) out->n_expanded[r] = (int32_t)t->size;
        if (out->n_pops) out->n_pops[r] = (int32_t)n_pops;
        if (out->pop_hash) out->pop_hash[r] = hash;
        for (int d = 0; d <= Hp; ++d) {
            size_t o = (size_t)r * (Hp + 1) + d;
            if (out->trims) out->trims[o] = exhausted ? (d == 0 ? in->trim0[r] : 0) : t->trim[path[d] * nV + v];
            if (out->tree_path) out->tree_path[o] = exhausted ? 0 : (int32_t)path[d];
            if (out->g_path) out->g_path[o] = exhausted ? NAN : t->g[path[d]];
            if (out->h_path) out->h_path[o] = exhausted ? NAN : t->h[path[d]];
        }
        for (int d = 1; d <= Hp; ++d) {
            size_t o = (size_t)r * Hp + (d - 1);
            if (out->y_predicted) {
                out->y_predicted[o * 3 + 0] = exhausted ? NAN : t->x[path[d] * nV + v];
                out->y_predicted[o * 3 + 1] = exhausted ? NAN : t->y[path[d] * nV + v];
                out->y_predicted[o * 3 + 2] = exhausted ? NAN : t->yaw[path[d] * nV + v];
            }
            if (out->shape_npts) {
                double px[PDMPC_AREA_STRIDE], py[PDMPC_AREA_STRIDE];
                int ns = 0;
                for (int i = 0; i < PDMPC_AREA_STRIDE; ++i) px[i] = py[i] = 0.0;
                if (!exhausted) {
                    int64_t pa = path[d - 1], ch = path[d];
                    int edge = edge_of[(t->trim[pa * nV + v] - 1) * nT + (t->trim[ch * nV + v] - 1)];
                    double c, sn2;
                    oracle_sincos(t->yaw[pa * nV + v], &sn2, &c);
                    place_area(mpa, edge, PDMPC_AREA_NORMAL, c, sn2, t->x[pa * nV + v], t->y[pa * nV + v], px, py, &ns);
                }
                out->shape_npts[o] = ns;
                if (out->shape_x && out->shape_y)
                    for (int i = 0; i < PDMPC_AREA_STRIDE; ++i) {
                        out->shape_x[o * PDMPC_AREA_STRIDE + i] = px[i];
                        out->shape_y[o * PDMPC_AREA_STRIDE + i] = py[i];
                    }
            }
        }
    }
    free(t->x); free(t->y); free(t->yaw); free(t->trim); free(t->g); free(t->h); free(t->k); free(t->parent);
    free(pq.a);
    return 0;
}

static int32_t *build_edge_of(const pdmpc_mpa_desc *mpa) {
    int nT = mpa->n_trims;
    int32_t *e = (int32_t *)malloc((size_t)nT * nT * sizeof(int32_t));
    for (int i = 0; i < nT * nT; ++i) e[i] = -1;
    for (int i = 0; i < mpa->n_edges; ++i)
        e[(mpa->edge_from[i] - 1) * nT + (mpa->edge_to[i] - 1)] = i;
    return e;
}

typedef struct {
    const pdmpc_mpa_desc *mpa;
    const pdmpc_batch_in *in;
    pdmpc_batch_out *out;
    int32_t *edge_of;
    int n;
    int *next; /* shared work counter: dynamic chunks, so a heavy search does not idle 7 cores */
} job_t;

static void *job_main(void *p) {
    job_t *j = (job_t *)p;
    work_t w;
    memset(&w, 0, sizeof(w));
    w.edge_of = j->edge_of;
    for (;;) {
        int b = __atomic_fetch_add(j->next, 16, __ATOMIC_RELAXED);
        if (b >= j->n) break;
        int e = b + 16 < j->n ? b + 16 : j->n;
        for (int si = b; si < e; ++si) search_one(j->mpa, j->in, j->out, si, &w, NULL);
    }
    tree_free(&w.tree);
    free(w.pq.a);
    free(w.ox);
    free(w.oy);
    return NULL;
}

int oracle_plan_batch(const pdmpc_mpa_desc *mpa, const pdmpc_batch_in *in, pdmpc_batch_out *out,
                      int n_threads) {
    if (!mpa || !in || !out) return PDMPC_ERR_BAD_INPUT;
    int n = in->n_searches;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n) n_threads = n > 0 ? n : 1;
    int32_t *edge_of = build_edge_of(mpa);
    int next = 0;
    job_t *jobs = (job_t *)calloc((size_t)n_threads, sizeof(job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].mpa = mpa; jobs[t].in = in; jobs[t].out = out; jobs[t].edge_of = edge_of;
        jobs[t].n = n;
        jobs[t].next = &next;
    }
    if (n_threads == 1) {
        job_main(&jobs[0]);
    } else {
        for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, job_main, &jobs[t]);
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    }
    free(jobs);
    free(th);
    free(edge_of);
    return PDMPC_OK;
}

int oracle_plan_trace(const pdmpc_mpa_desc *mpa, const pdmpc_batch_in *in, int trace_search,
                      int64_t *pop_trace, int64_t trace_cap, int64_t *n_trace) {
    if (!mpa || !in || trace_search < 0 || trace_search >= in->n_searches) return PDMPC_ERR_BAD_INPUT;
    work_t w;
    memset(&w, 0, sizeof(w));
    w.edge_of = build_edge_of(mpa);
    trace_t tr = {pop_trace, trace_cap, 0};
    pdmpc_batch_out out;
    memset(&out, 0, sizeof(out));
    search_one(mpa, in, &out, trace_search, &w, &tr);
    if (n_trace) *n_trace = tr.n;
    tree_free(&w.tree);
    free(w.pq.a);
    free(w.ox);
    free(w.oy);
    free(w.edge_of);
    return PDMPC_OK;
}

/* ========================================================================= */
/* MonteCarloTreeSearch (OptimizerType.MatlabSampled), SURVEY.md §8 a10.      */
/* hlc/optimizer/graph_search/MonteCarloTreeSearch.m:29-251                   */

/* MATLAB RandStream('mt19937ar', Seed = s) + rand: the published MT19937
 * (Matsumoto & Nishimura, mt19937ar.c: init_genrand, genrand_int32,
 * genrand_res53).  Third-party arithmetic that is not under /root/reference;
 * pinned in tests/ against numpy.random.RandomState (the same published
 * algorithm, independent implementation).  MATLAB maps Seed = 0 to the
 * generator's default seed 5489 (unpinned; time_step + vehicle_index >= 2 in
 * the reference's only call site, MonteCarloTreeSearch.m:31). */
typedef struct {
    uint32_t mt[624];
    int idx;
} mt_t;

static void mt_seed(mt_t *m, uint32_t s) {
    if (s == 0) s = 5489u;
    m->mt[0] = s;
    for (int i = 1; i < 624; ++i)
        m->mt[i] = 1812433253u * (m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) + (uint32_t)i;
    m->idx = 624;
}

static uint32_t mt_next(mt_t *m) {
    if (m->idx >= 624) {
        uint32_t *mt = m->mt;
        for (int kk = 0; kk < 624; ++kk) {
            uint32_t y = (mt[kk] & 0x80000000u) | (mt[(kk + 1) % 624] & 0x7fffffffu);
            mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->idx = 0;
    }
    uint32_t y = m->mt[m->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

static double mt_rand(mt_t *m) { /* genrand_res53 */
    uint32_t a = mt_next(m) >> 5, b = mt_next(m) >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}

void oracle_mt19937_rand(uint32_t seed, double *out, int n) {
    mt_t m;
    mt_seed(&m, seed);
    for (int i = 0; i < n; ++i) out[i] = mt_rand(&m);
}

/* successor_trims{t, k} = find(transition_matrix_single(t, :, k))
 * (MotionPrimitiveAutomaton.m:150); returns the count, fills succ[] 1-based trims */
static int successors_of(const pdmpc_mpa_desc *mpa, int trim, int k, int *succ) {
    const int nT = mpa->n_trims;
    const uint8_t *row = mpa->transition + ((size_t)(k - 1) * nT + (trim - 1)) * nT;
    int n = 0;
    for (int j = 0; j < nT; ++j)
        if (row[j]) succ[n++] = j + 1;
    return n;
}

/* MonteCarloTreeSearch.do_graph_search (:40-251) for search `si`. */
static int mcts_one(const pdmpc_mpa_desc *mpa, const pdmpc_batch_in *in, const pdmpc_mcts_params *prm,
                    pdmpc_batch_out *out, int si, work_t *w) {
    const int Hp = mpa->Hp, nT = mpa->n_trims;
    const int nmax = prm->n_expansions_max;
    int status = PDMPC_OK;
    /* :53 maximum_branching_factor = max(sum(transition_matrix_single, 2), [], 'all') */
    int B = 0;
    for (size_t r = 0; r < (size_t)Hp * nT; ++r) {
        int c = 0;
        for (int j = 0; j < nT; ++j) c += mpa->transition[r * nT + j] != 0;
        if (c > B) B = c;
    }
    /* :31,52 */
    const int n_rand = Hp * nmax;
    double *rnd = (double *)malloc((size_t)(n_rand > 0 ? n_rand : 1) * sizeof(double));
    {
        mt_t m;
        mt_seed(&m, prm->seed[si]);
        for (int i = 0; i < n_rand; ++i) rnd[i] = mt_rand(&m);
    }
    /* :59-70 (MATLAB grows the arrays on assignment).  The budget is only tested between
     * roll-outs (:86), so the last roll-out may overshoot: n_expansions <= nmax + Hp - 1 */
    const int cap = nmax + Hp + 2;
    int32_t *trims = (int32_t *)calloc((size_t)cap, sizeof(int32_t));
    int32_t *parents = (int32_t *)calloc((size_t)cap, sizeof(int32_t));
    int32_t *children = (int32_t *)calloc((size_t)cap * (B > 0 ? B : 1), sizeof(int32_t)); /* [node][position] */
    double *shx = (double *)calloc((size_t)cap * PDMPC_AREA_STRIDE, sizeof(double));        /* shapes_tmp */
    double *shy = (double *)calloc((size_t)cap * PDMPC_AREA_STRIDE, sizeof(double));
    int32_t *shn = (int32_t *)calloc((size_t)cap, sizeof(int32_t));
    int *succ = (int *)malloc((size_t)nT * sizeof(int));
    const double root_pose[3] = {in->x0[si], in->y0[si], in->yaw0[si]};
    trims[1] = in->trim0[si];
    parents[1] = 0;
    {
        int n = successors_of(mpa, trims[1], 1, succ);
        for (int i = 0; i < n; ++i) children[1 * B + i] = 1;
    }
    int n_nodes = 1;
    oracle_pq *pq = &w->pq; /* valid_nodes_at_hp */
    pq->len = 0;

    /* :75-76 set_up_constraints: lanelet = [left, NaN, right, NaN] */
    int l0 = in->lane_ptr[2 * si], l1 = in->lane_ptr[2 * si + 1], l2 = in->lane_ptr[2 * si + 2];
    int nlane = (l2 - l0) + 2;
    double *lanex = (double *)malloc((size_t)nlane * sizeof(double));
    double *laney = (double *)malloc((size_t)nlane * sizeof(double));
    {
        int n = 0;
        for (int i = l0; i < l1; ++i) { lanex[n] = in->lane_x[i]; laney[n] = in->lane_y[i]; ++n; }
        lanex[n] = NAN; laney[n] = NAN; ++n;
        for (int i = l1; i < l2; ++i) { lanex[n] = in->lane_x[i]; laney[n] = in->lane_y[i]; ++n; }
        lanex[n] = NAN; laney[n] = NAN; ++n;
    }
    const double *rx = in->ref_x + (size_t)si * Hp, *ry = in->ref_y + (size_t)si * Hp;

    int n_expansions = 0, n_traversals = 0, is_finished = 0;
    uint64_t hash = 0xcbf29ce484222325ULL;

    while (n_expansions < nmax && !is_finished && status == PDMPC_OK) { /* :86 */
        int node_id = 1;
        double solution_cost = 0;
        double node_pose[3] = {root_pose[0], root_pose[1], root_pose[2]};
        int is_valid = 0, child_position = 0, node_parent = 0;
        for (int i_step = 1; i_step <= Hp; ++i_step) { /* :92 */
            is_valid = 0;
            n_traversals = n_traversals + 1;
            /* :97-98 trim_positions = find(children(:, node_id)) */
            int n_trims = 0;
            for (int p = 0; p < B; ++p) n_trims += children[node_id * B + p] != 0;
            if (n_trims != 0) {
                if (n_traversals > n_rand) { status = PDMPC_ERR_CAPACITY; break; } /* MATLAB: index out of bounds */
                /* :102 trim_positions(ceil(r * n_trims)) */
                int pick = (int)ceil(rnd[n_traversals - 1] * (double)n_trims);
                int seen = 0;
                for (int p = 0; p < B; ++p)
                    if (children[node_id * B + p] != 0 && ++seen == pick) { child_position = p + 1; break; }
            } else {
                if (node_id != 1) { /* :106-109 remove edge to node without children */
                    int parent_id = parents[node_id];
                    for (int p = 0; p < B; ++p)
                        if (children[parent_id * B + p] == node_id) children[parent_id * B + p] = 0;
                    break;
                } else { /* :110-113 */
                    is_finished = 1;
                    break;
                }
            }
            hash = fnv1a_u32(hash, ((uint32_t)node_id << 8) | (uint32_t)child_position);
            /* :117-122 */
            int parent_trim = trims[node_id];
            successors_of(mpa, parent_trim, i_step, succ);
            int goal_trim = succ[child_position - 1];
            int edge = w->edge_of[(parent_trim - 1) * nT + (goal_trim - 1)];
            double c, s;
            oracle_sincos(node_pose[2], &s, &c); /* :124-125 */
            double start_pose[3] = {node_pose[0], node_pose[1], node_pose[2]};
            /* :127-132 node_pose + [c -s 0; s c 0; 0 0 1] * dpose (the zero products add exact zeros) */
            double dx = mpa->edge_dx[edge], dy = mpa->edge_dy[edge], dyaw = mpa->edge_dyaw[edge];
            node_pose[0] = node_pose[0] + (c * dx - s * dy);
            node_pose[1] = node_pose[1] + (s * dx + c * dy);
            node_pose[2] = node_pose[2] + dyaw;
            /* :137 */
            {
                double ddx = node_pose[0] - rx[i_step - 1], ddy = node_pose[1] - ry[i_step - 1];
                double nrm = sqrt(ddx * ddx + ddy * ddy);
                solution_cost = solution_cost + nrm * nrm;
            }
            /* :139-144 */
            if (children[node_id * B + child_position - 1] != 1) {
                node_id = children[node_id * B + child_position - 1];
                continue;
            }
            n_expansions = n_expansions + 1; /* :146 */
            node_parent = node_id;
            /* :150-159 */
            double sx[PDMPC_AREA_STRIDE], sy[PDMPC_AREA_STRIDE], bx[PDMPC_AREA_STRIDE], by[PDMPC_AREA_STRIDE];
            int ns, nb, n_child_succ = 0;
            place_area(mpa, edge, PDMPC_AREA_NORMAL, c, s, start_pose[0], start_pose[1], sx, sy, &ns);
            if (i_step != Hp) {
                place_area(mpa, edge, PDMPC_AREA_WITHOUT_OFFSET, c, s, start_pose[0], start_pose[1], bx, by, &nb);
                n_child_succ = successors_of(mpa, goal_trim, i_step + 1, succ);
            } else {
                place_area(mpa, edge, PDMPC_AREA_LARGE_OFFSET, c, s, start_pose[0], start_pose[1], bx, by, &nb);
            }
            /* :161-170 */
            if (in->checker == PDMPC_CHECKER_SAT)
                is_valid = constraints_sat(in, Hp, si, i_step, sx, sy, ns, bx, by, nb);
            else
                is_valid = constraints_interx(in, Hp, si, i_step, w, sx, sy, ns, bx, by, nb, lanex, laney, nlane);
            if (!is_valid) { /* :172-175 remove edge */
                children[node_parent * B + child_position - 1] = 0;
                break;
            } else { /* :176-185 add node */
                n_nodes = n_nodes + 1;
                parents[n_nodes] = node_parent;
                trims[n_nodes] = goal_trim;
                for (int i = 0; i < n_child_succ; ++i) children[n_nodes * B + i] = 1;
                children[node_parent * B + child_position - 1] = n_nodes;
                shn[n_nodes] = ns;
                for (int i = 0; i < ns; ++i) {
                    shx[(size_t)n_nodes * PDMPC_AREA_STRIDE + i] = sx[i];
                    shy[(size_t)n_nodes * PDMPC_AREA_STRIDE + i] = sy[i];
                }
                node_id = n_nodes;
            }
        }
        if (is_valid) { /* :189-193 */
            oracle_pq_push(pq, node_id, solution_cost);
            children[node_parent * B + child_position - 1] = 0; /* avoid double exploration */
        }
    }

    double cost = 0;
    int64_t best = oracle_pq_pop(pq, &cost); /* :197 */
    int exhausted = (best == -1) || status != PDMPC_OK; /* :202-205 */

    if (out->status) out->status[si] = status;
    if (out->is_exhausted) out->is_exhausted[si] = (uint8_t)exhausted;
    if (out->n_expanded) out->n_expanded[si] = n_expansions; /* :199 */
    if (out->n_pops) out->n_pops[si] = n_traversals;
    if (out->pop_hash) out->pop_hash[si] = hash;
    int64_t path[PDMPC_MAX_HP + 1];
    double pose[PDMPC_MAX_HP + 1][3];
    if (!exhausted) { /* :215-232 */
        int64_t n = best;
        for (int d = Hp; d >= 0; --d) {
            path[d] = n;
            n = parents[n];
        }
        pose[0][0] = root_pose[0]; pose[0][1] = root_pose[1]; pose[0][2] = root_pose[2];
        for (int i = 1; i <= Hp; ++i) {
            int edge = w->edge_of[(trims[path[i - 1]] - 1) * nT + (trims[path[i]] - 1)];
            double c, s;
            oracle_sincos(pose[i - 1][2], &s, &c);
            double dx = mpa->edge_dx[edge], dy = mpa->edge_dy[edge], dyaw = mpa->edge_dyaw[edge];
            pose[i][0] = pose[i - 1][0] + (c * dx - s * dy);
            pose[i][1] = pose[i - 1][1] + (s * dx + c * dy);
            pose[i][2] = pose[i - 1][2] + dyaw;
        }
    }
    for (int d = 0; d <= Hp; ++d) {
        size_t o = (size_t)si * (Hp + 1) + d;
        if (out->trims) out->trims[o] = exhausted ? (d == 0 ? in->trim0[si] : 0) : trims[path[d]];
        if (out->tree_path) out->tree_path[o] = exhausted ? 0 : (int32_t)path[d];
        /* :211-213,239 tree.g = -1 except the best leaf, tree.h = -1 */
        if (out->g_path) out->g_path[o] = exhausted ? NAN : (d == Hp ? cost : -1.0);
        if (out->h_path) out->h_path[o] = exhausted ? NAN : -1.0;
    }
    for (int d = 1; d <= Hp; ++d) {
        size_t o = (size_t)si * Hp + (d - 1);
        if (out->y_predicted) /* :242 */
            for (int c3 = 0; c3 < 3; ++c3) out->y_predicted[o * 3 + c3] = exhausted ? NAN : pose[d][c3];
        if (out->shape_npts) { /* :243 return_path_area(shapes_tmp, ...) */
            out->shape_npts[o] = exhausted ? 0 : shn[path[d]];
            if (out->shape_x && out->shape_y)
                for (int i = 0; i < PDMPC_AREA_STRIDE; ++i) {
                    out->shape_x[o * PDMPC_AREA_STRIDE + i] =
                        exhausted ? 0.0 : shx[(size_t)path[d] * PDMPC_AREA_STRIDE + i];
                    out->shape_y[o * PDMPC_AREA_STRIDE + i] =
                        exhausted ? 0.0 : shy[(size_t)path[d] * PDMPC_AREA_STRIDE + i];
                }
        }
    }
    free(rnd); free(trims); free(parents); free(children); free(shx); free(shy); free(shn); free(succ);
    free(lanex); free(laney);
    return 0;
}

typedef struct {
    const pdmpc_mpa_desc *mpa;
    const pdmpc_batch_in *in;
    const pdmpc_mcts_params *prm;
    pdmpc_batch_out *out;
    int32_t *edge_of;
    int n;
    int *next;
} mjob_t;

static void *mjob_main(void *p) {
    mjob_t *j = (mjob_t *)p;
    work_t w;
    memset(&w, 0, sizeof(w));
    w.edge_of = j->edge_of;
    for (;;) {
        int b = __atomic_fetch_add(j->next, 16, __ATOMIC_RELAXED);
        if (b >= j->n) break;
        int e = b + 16 < j->n ? b + 16 : j->n;
        for (int si = b; si < e; ++si) mcts_one(j->mpa, j->in, j->prm, j->out, si, &w);
    }
    free(w.pq.a);
    free(w.ox);
    free(w.oy);
    return NULL;
}

int oracle_mcts_plan_batch(const pdmpc_mpa_desc *mpa, const pdmpc_batch_in *in, const pdmpc_mcts_params *prm,
                           pdmpc_batch_out *out, int n_threads) {
    if (!mpa || !in || !prm || !out || !prm->seed || prm->n_expansions_max < 1) return PDMPC_ERR_BAD_INPUT;
    int n = in->n_searches;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n) n_threads = n > 0 ? n : 1;
    int32_t *edge_of = build_edge_of(mpa);
    int next = 0;
    mjob_t *jobs = (mjob_t *)calloc((size_t)n_threads, sizeof(mjob_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].mpa = mpa; jobs[t].in = in; jobs[t].prm = prm; jobs[t].out = out;
        jobs[t].edge_of = edge_of; jobs[t].n = n; jobs[t].next = &next;
    }
    if (n_threads == 1) {
        mjob_main(&jobs[0]);
    } else {
        for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, mjob_main, &jobs[t]);
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    }
    free(jobs);
    free(th);
    free(edge_of);
    return PDMPC_OK;
}


/* Joint searches of the batch, one after the other (rows = searches x n_vehicles). */
int oracle_joint_plan_batch(const pdmpc_mpa_desc *mpa, const pdmpc_batch_in *in, int n_vehicles,
                            pdmpc_batch_out *out, int64_t max_nodes) {
    if (n_vehicles < 1 || n_vehicles > PDMPC_MAX_JOINT || in->n_searches % n_vehicles) return PDMPC_ERR_BAD_INPUT;
    int32_t *edge_of = build_edge_of(mpa);
    for (int s = 0; s < in->n_searches / n_vehicles; ++s)
        joint_search_one(mpa, in, out, s, n_vehicles, edge_of, max_nodes);
    free(edge_of);
    return PDMPC_OK;
}
