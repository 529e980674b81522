/*
 * pdmpc_oracle.h — CPU oracle for the p-dmpc graph-search optimizer.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (MATLAB + one C++ MEX) used as the parity checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
 * Nothing under p-dmpc_b200/ may include, link or call it.
 *
 * Pinning status (see oracle/README.md):
 *   - intersect_sat / intersect_lanelets: pinned by the reference's own KATs
 *     (tests/unittests/hlc/intersect_unittest.m:8-54).
 *   - priority-queue pop order: pinned against the reference's unmodified
 *     priority_queue_interface_mex.cpp compiled here against stub MEX headers
 *     (oracle/_ref/libpq_ref.so) and this container's libstdc++.
 *   - search loop, expand_node, InterX: the reference holds no golden vectors
 *     for them and MATLAB cannot run here: PARITY UNPINNED
 *     beyond the two items above; restated line by line with citations.
 *
 * It shares only the boundary structs with the product (include/pdmpc_b200.h).
 */
#ifndef PDMPC_ORACLE_H
#define PDMPC_ORACLE_H

#include "../include/pdmpc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* sin/cos as specified in DESIGN.md §"sincos" (Cody-Waite + Taylor, no FMA). */
void oracle_sincos(double x, double *s, double *c);

/* intersect_sat.m:1-42.  Polygons as SoA, n points each (open or closed). */
int oracle_intersect_sat(const double *x1, const double *y1, int n1,
                         const double *x2, const double *y2, int n2);

/* intersect_lanelets.m:1-22 (only the reference's unit tests call it).
 * lanelet columns as separate arrays, n rows. */
int oracle_intersect_lanelets(const double *sx, const double *sy, int ns,
                              const double *rx, const double *ry,
                              const double *lx, const double *ly, int n);

/* intersect_lanelet_boundary.m:1-56 */
int oracle_intersect_lanelet_boundary(const double *sx, const double *sy, int ns,
                                      const double *lx, const double *ly, int nl,
                                      const double *rx, const double *ry, int nr);

/* InterX.m:48-85,108-110 with isReturnPoints = false.  NaN columns allowed. */
int oracle_interx(const double *x1, const double *y1, int n1,
                  const double *x2, const double *y2, int n2);

/* libstdc++ std::priority_queue<(id,value), vector, value-greater> restated
 * (stl_heap.h __push_heap / __adjust_heap). */
typedef struct oracle_pq oracle_pq;
oracle_pq *oracle_pq_new(void);
void oracle_pq_free(oracle_pq *q);
void oracle_pq_push(oracle_pq *q, int64_t id, double value);
int64_t oracle_pq_pop(oracle_pq *q, double *value); /* -1 when empty */
int64_t oracle_pq_size(const oracle_pq *q);
/* diagnostic: number of pops (since the last reset) whose minimum was not unique */
int64_t oracle_tie_pops(int reset);

/* GraphSearch.do_graph_search for every search of the batch, n_threads host
 * threads (dynamic chunks of 16 searches).  Optional trace: if pop_trace != NULL it
 * receives the popped node ids of search `trace_search` (up to trace_cap). */
int oracle_plan_batch(const pdmpc_mpa_desc *mpa, const pdmpc_batch_in *in,
                      pdmpc_batch_out *out, int n_threads);
int oracle_plan_trace(const pdmpc_mpa_desc *mpa, const pdmpc_batch_in *in,
                      int trace_search, int64_t *pop_trace, int64_t trace_cap,
                      int64_t *n_trace);

/* Centralized (joint) search, iter.amount = n_vehicles > 1: rows of the batch are searches x vehicles, the
 * obstacle slots of a search are those of its first row; SAT checker (CentralizedController.m:33-59,
 * expand_node.m:15-75, are_constraints_satisfied_sat.m:15-53).  max_nodes: node capacity as on the device. */
int oracle_joint_plan_batch(const pdmpc_mpa_desc *mpa, const pdmpc_batch_in *in, int n_vehicles,
                            pdmpc_batch_out *out, int64_t max_nodes);

/* pop_hash over the popped nodes that passed their edge check only (CUDA launch shape 5); default off */
void oracle_set_hash_valid_pops_only(int on);

/* MonteCarloTreeSearch.do_graph_search (MonteCarloTreeSearch.m:40-251) for every search. */
int oracle_mcts_plan_batch(const pdmpc_mpa_desc *mpa, const pdmpc_batch_in *in,
                           const pdmpc_mcts_params *prm, pdmpc_batch_out *out, int n_threads);
/* rand(RandStream('mt19937ar', Seed = seed), 1, n): MT19937 + genrand_res53 */
void oracle_mt19937_rand(uint32_t seed, double *out, int n);

#ifdef __cplusplus
}
#endif
#endif
