/* pcp_internal.h -- private interface between pcp_engine.cu and pcp_search.cpp (same .so).
 * Not part of the C ABI of include/pcp_b200.h. */
#ifndef PCP_INTERNAL_H
#define PCP_INTERNAL_H
#include <stdint.h>
#include "../../include/pcp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pcp_burst_result {
  int32_t status;  /* why the last burst ended: 0 slice used up, 1 solution (one-solution mode),
                      -1 tree exhausted (one-solution mode), 2 end of search */
  int32_t err;
  uint64_t nodes, solutions, failures, iterations, propagations; /* cumulative since begin */
  double kernel_seconds;                                         /* cumulative device time */
} pcp_burst_result;

/* Device-resident DFS: OneSolution/AllSolution o StopNode o Propagation o
 * Brancher(FirstSmallestVar, MiddleVal, BinarySplit) run in bursts of nodes inside one launch.
 * begin: the current engine state becomes the root; end: the engine is restored to it. */
int pcp_internal_burst_supported(pcp_engine* e, const pcp_search_config* cfg, uint64_t trace_capacity);
int pcp_internal_burst_begin(pcp_engine* e, int32_t all_solutions, uint64_t node_limit, uint64_t trace_capacity,
                             int32_t trace_domains);
int pcp_internal_burst_step(pcp_engine* e, uint64_t max_nodes, pcp_burst_result* res);
/* the next slice of n device searches in one launch (a group of CTAs per search); *fused = 0 and
 * nothing done when their engines cannot share a launch */
int pcp_internal_burst_step_many(pcp_engine* const* es, int32_t n, const uint64_t* budgets, pcp_burst_result* res, int32_t* fused);
/* copies the per-node trace of nodes [first, first+n): status and (if recorded) domains */
int pcp_internal_burst_trace(pcp_engine* e, uint64_t first, uint64_t n, int32_t* status, int32_t* lo, int32_t* hi);
/* IntervalSet engines: the value sets of traced nodes as bit windows [n * V, words] from `base` */
int pcp_internal_burst_trace_bits(pcp_engine* e, uint64_t first, uint64_t n, const int32_t* lo, const int32_t* hi, int32_t base,
                                  int32_t words, uint32_t* out);
int pcp_internal_burst_end(pcp_engine* e);
/* split-phase pcp_consistency: launch / non-blocking "result there?" / collect */
int pcp_internal_consistency_begin(pcp_engine* e);
int pcp_internal_consistency_poll(pcp_engine* e);
int pcp_internal_consistency_end(pcp_engine* e, int32_t* status, pcp_stats* stats);
/* 1 when the engine was created with PCP_FLAG_INTERVAL_SET */
int pcp_internal_interval_set(const pcp_engine* e);

#ifdef __cplusplus
}
#endif
#endif
