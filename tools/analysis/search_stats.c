/* search_stats.c — offline instrumentation of the search on recorded batches (design aid, not product,
 * not parity infrastructure).  Includes the oracle restatement and re-runs its loop with counters:
 *   - valid / invalid pops, queue length at pop, ties
 *   - InterX filter selectivity: segments whose line can cross the shape's bounding box (monotone-corner
 *     test on the reference's own C2 expression) and C1&C2 survivors
 *   - front-buffer queue policy (K sorted entries in registers + backing heap): pops served by the front
 * build: gcc -O2 -ffp-contract=off -shared -fPIC -o build/libstats.so tools/analysis/search_stats.c -lm -lpthread
 */
#include <stdio.h>
#include "../../oracle/pdmpc_oracle.c"

typedef struct {
    int64_t pops, valid_pops, goal_pops, expansions, nodes, ties;
    int64_t qlen_hist[16];         /* log2 buckets of queue length at pop */
    int64_t segs, segs_bbox, segs_c2row, pairs_hit;   /* InterX */
    int64_t checks, checks_any_bbox;
    int64_t front_pops, back_pops, back_pushes, front_inserts, evictions;
    int64_t child_is_next;         /* pop returns a child of the previous expansion */
    int64_t invalid_children, children;
} stats_t;

static int K_FRONT = 32;
void stats_set_front(int k) { K_FRONT = k; }

/* monotone-corner filter on b(V) = (Vy*dx2 - Vx*dy2) - S2 over the shape's bounding box */
static void interx_stats(const double *x1, const double *y1, int n1, const double *x2, const double *y2, int n2,
                         stats_t *st) {
    if (n1 == 0 || n2 == 0) return;
    double xl = x1[0], xh = x1[0], yl = y1[0], yh = y1[0];
    for (int i = 1; i < n1; ++i) {
        xl = fmin(xl, x1[i]); xh = fmax(xh, x1[i]); yl = fmin(yl, y1[i]); yh = fmax(yh, y1[i]);
    }
    int any = 0;
    for (int j = 0; j + 1 < n2; ++j) {
        ++st->segs;
        double dx2 = x2[j + 1] - x2[j], dy2 = y2[j + 1] - y2[j];
        double S2 = dx2 * y2[j] - dy2 * x2[j];
        double pyl = yl * dx2, pyh = yh * dx2, qxl = xl * dy2, qxh = xh * dy2;
        double bmax = (fmax(pyl, pyh) - fmin(qxl, qxh)) - S2;
        double bmin = (fmin(pyl, pyh) - fmax(qxl, qxh)) - S2;
        int excluded = (bmin > 0) || (bmax < 0) || (dx2 != dx2) || (dy2 != dy2);
        /* exact rows */
        int c2row = 0, hit = 0;
        for (int i = 0; i + 1 < n1; ++i) {
            double b0 = (y1[i] * dx2 - x1[i] * dy2) - S2;
            double b1 = (y1[i + 1] * dx2 - x1[i + 1] * dy2) - S2;
            int c2 = (b0 * b1) < 0;
            if (c2) {
                c2row = 1;
                double dx1 = x1[i + 1] - x1[i], dy1 = y1[i + 1] - y1[i];
                double S1 = dx1 * y1[i] - dy1 * x1[i];
                double a0 = (dx1 * y2[j] - dy1 * x2[j]) - S1;
                double a1 = (dx1 * y2[j + 1] - dy1 * x2[j + 1]) - S1;
                if ((a0 * a1) < 0) hit = 1;
            }
        }
        if (excluded && c2row) { fprintf(stderr, "FILTER VIOLATION\n"); abort(); }
        if (!excluded) { ++st->segs_bbox; any = 1; }
        if (c2row) ++st->segs_c2row;
        if (hit) ++st->pairs_hit;
    }
    ++st->checks;
    if (any) ++st->checks_any_bbox;
}

/* front buffer simulation: sorted array of at most K entries, back = multiset size counter + min tracking via heap */
typedef struct { double f; int64_t id; } fent;
typedef struct {
    fent front[64]; int nf;
    oracle_pq back;
} fq_t;
static void fq_push(fq_t *q, int64_t id, double f, stats_t *st) {
    double back_min = q->back.len ? q->back.a[0].val : INFINITY;
    int full = q->nf >= K_FRONT;
    if ((q->nf > 0 && f < q->front[q->nf - 1].f) || (!full && f <= back_min)) {
        /* insert sorted (after equal entries) */
        int p = q->nf;
        while (p > 0 && q->front[p - 1].f > f) --p;
        if (full) {
            fent ev = q->front[q->nf - 1];
            oracle_pq_push(&q->back, ev.id, ev.f); ++st->back_pushes; ++st->evictions;
            --q->nf;
            if (p > q->nf) p = q->nf;
        }
        for (int i = q->nf; i > p; --i) q->front[i] = q->front[i - 1];
        q->front[p].f = f; q->front[p].id = id; ++q->nf; ++st->front_inserts;
    } else {
        oracle_pq_push(&q->back, id, f); ++st->back_pushes;
    }
}
static int64_t fq_pop(fq_t *q, stats_t *st) {
    if (q->nf > 0) {
        int64_t id = q->front[0].id;
        for (int i = 1; i < q->nf; ++i) q->front[i - 1] = q->front[i];
        --q->nf; ++st->front_pops;
        return id;
    }
    if (q->back.len == 0) return -1;
    ++st->back_pops;
    return oracle_pq_pop(&q->back, NULL);
}

static void stats_one(const pdmpc_mpa_desc *mpa, const pdmpc_batch_in *in, int si, work_t *w, stats_t *st) {
    const int Hp = mpa->Hp, nT = mpa->n_trims;
    tree_t *t = &w->tree;
    oracle_pq *pq = &w->pq;
    pq->len = 0; t->size = 0; tree_reserve(t, 1);
    t->size = 1; t->x[1] = in->x0[si]; t->y[1] = in->y0[si]; t->yaw[1] = in->yaw0[si];
    t->trim[1] = in->trim0[si]; t->k[1] = 0; t->g[1] = 0; t->h[1] = 0; t->parent[1] = 0;
    oracle_pq_push(pq, 1, 0.0);
    fq_t fq; memset(&fq, 0, sizeof fq);
    fq_push(&fq, 1, 0.0, st);
    int l0 = in->lane_ptr[2 * si], l1 = in->lane_ptr[2 * si + 1], l2 = in->lane_ptr[2 * si + 2];
    int nlane = (l2 - l0) + 2;
    double *lanex = (double *)malloc((size_t)nlane * sizeof(double));
    double *laney = (double *)malloc((size_t)nlane * sizeof(double));
    { int n = 0;
      for (int i = l0; i < l1; ++i) { lanex[n] = in->lane_x[i]; laney[n] = in->lane_y[i]; ++n; }
      lanex[n] = NAN; laney[n] = NAN; ++n;
      for (int i = l1; i < l2; ++i) { lanex[n] = in->lane_x[i]; laney[n] = in->lane_y[i]; ++n; }
      lanex[n] = NAN; laney[n] = NAN; ++n; }
    const double *rx = in->ref_x + (size_t)si * Hp, *ry = in->ref_y + (size_t)si * Hp;
    const double *vr = in->v_ref + (size_t)si * Hp;
    int64_t last_first = 0, last_last = -1;
    for (;;) {
        int64_t qlen = pq->len;
        int tie = 0;
        if (qlen >= 2) { double top = pq->a[0].val; if (!(top < pq->a[1].val) || (qlen > 2 && !(top < pq->a[2].val))) tie = 1; }
        int64_t id = oracle_pq_pop(pq, NULL);
        int64_t id2 = fq_pop(&fq, st);
        if (id == -1) break;
        if (!tie && id2 != id) { fprintf(stderr, "front queue mismatch without tie\n"); abort(); }
        if (tie) { ++st->ties; /* resync the front queue: rebuild from the exact heap */
            fq.nf = 0; fq.back.len = 0;
            if (id2 != id) { /* put everything of the exact heap into the sim */ }
            for (int64_t i = 0; i < pq->len; ++i) fq_push(&fq, pq->a[i].id, pq->a[i].val, st);
        }
        ++st->pops;
        { int b = 0; int64_t q = qlen; while (q > 1 && b < 15) { q >>= 1; ++b; } ++st->qlen_hist[b]; }
        if (id >= last_first && id <= last_last) ++st->child_is_next;
        int is_valid = 1;
        int32_t par = t->parent[id];
        if (par) {
            double pX = t->x[par], pY = t->y[par], pYaw = t->yaw[par];
            int t1 = t->trim[par], t2 = t->trim[id], cK = t->k[id];
            int edge = w->edge_of[(t1 - 1) * nT + (t2 - 1)];
            double c, s; oracle_sincos(pYaw, &s, &c);
            double sx[PDMPC_AREA_STRIDE], sy[PDMPC_AREA_STRIDE], bx[PDMPC_AREA_STRIDE], by[PDMPC_AREA_STRIDE];
            int ns, nb;
            place_area(mpa, edge, PDMPC_AREA_NORMAL, c, s, pX, pY, sx, sy, &ns);
            place_area(mpa, edge, cK == Hp ? PDMPC_AREA_LARGE_OFFSET : PDMPC_AREA_WITHOUT_OFFSET, c, s, pX, pY, bx, by, &nb);
            int64_t nobs = vectorize_step(in, Hp, si, cK, w);
            interx_stats(sx, sy, ns, w->ox, w->oy, (int)nobs, st);
            interx_stats(bx, by, nb, lanex, laney, nlane, st);
            is_valid = constraints_interx(in, Hp, si, cK, w, sx, sy, ns, bx, by, nb, lanex, laney, nlane);
        }
        if (!is_valid) continue;
        ++st->valid_pops;
        if (t->k[id] == Hp) { ++st->goal_pops; break; }
        ++st->expansions;
        {
            double curX = t->x[id], curY = t->y[id], curYaw = t->yaw[id], curG = t->g[id];
            int curTrim = t->trim[id];
            int k_exp = t->k[id] + 1;
            const uint8_t *row = mpa->transition + ((size_t)(k_exp - 1) * nT + (curTrim - 1)) * nT;
            int time_steps_to_go = Hp - k_exp;
            double c, s; oracle_sincos(curYaw, &s, &c);
            int64_t first_new = t->size + 1;
            for (int j = 0; j < nT; ++j) {
                if (!row[j]) continue;
                int edge = w->edge_of[(curTrim - 1) * nT + j];
                double dx = mpa->edge_dx[edge], dy = mpa->edge_dy[edge], dyaw = mpa->edge_dyaw[edge];
                double ex = c * dx - s * dy + curX, ey = s * dx + c * dy + curY, eyaw = curYaw + dyaw;
                double ddx = ex - rx[k_exp - 1], ddy = ey - ry[k_exp - 1];
                double nrm = sqrt(ddx * ddx + ddy * ddy);
                double eg = curG + nrm * nrm;
                double eh = 0, d_traveled_max = 0;
                for (int it = 1; it <= time_steps_to_go; ++it) {
                    d_traveled_max = d_traveled_max + in->dt_seconds * vr[k_exp + it - 1];
                    double hx = ex - rx[k_exp + it - 1], hy = ey - ry[k_exp + it - 1];
                    double hn = sqrt(hx * hx + hy * hy);
                    double m = fmax(0.0, hn - d_traveled_max);
                    eh = eh + m * m;
                }
                tree_reserve(t, t->size + 1);
                int64_t nid = ++t->size;
                t->x[nid] = ex; t->y[nid] = ey; t->yaw[nid] = eyaw; t->trim[nid] = j + 1; t->k[nid] = k_exp;
                t->g[nid] = eg; t->h[nid] = eh; t->parent[nid] = (int32_t)id;
            }
            last_first = first_new; last_last = t->size;
            for (int64_t nid = first_new; nid <= t->size; ++nid) {
                oracle_pq_push(pq, nid, t->g[nid] * 1 + t->h[nid] * 1);
                fq_push(&fq, nid, t->g[nid] * 1 + t->h[nid] * 1, st);
                ++st->children;
            }
        }
    }
    st->nodes += t->size;
    free(lanex); free(laney); free(fq.back.a);
}

int stats_batch(const pdmpc_mpa_desc *mpa, const pdmpc_batch_in *in, int first, int count, int64_t *out, int n_out) {
    work_t w; memset(&w, 0, sizeof w);
    w.edge_of = build_edge_of(mpa);
    stats_t st; memset(&st, 0, sizeof st);
    for (int si = first; si < first + count && si < in->n_searches; ++si) stats_one(mpa, in, si, &w, &st);
    int64_t *p = (int64_t *)&st;
    for (int i = 0; i < n_out && i < (int)(sizeof st / 8); ++i) out[i] = p[i];
    tree_free(&w.tree); free(w.pq.a); free(w.ox); free(w.oy); free(w.edge_of);
    return (int)(sizeof st / 8);
}
