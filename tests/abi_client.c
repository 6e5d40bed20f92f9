/* abi_client.c -- a plain-C client of include/pcp_b200.h (no Python, no C++): what a host in
 * another language sees of the boundary.  Built by __graft_entry__.build() into tests/abi_client,
 * run by tests/test_abi.py on the GPU box.
 *
 * It drives example/src/nqueens.rs for N = 8 to its 92 solutions (search/engine/all_solution.rs:
 * 67-74) three ways:
 *   1. the host-driven node loop a libpcp host runs, written out against the store surface:
 *      Branch::commit (pcp_restore + pcp_prop_alloc), Propagation::enter (pcp_consistency),
 *      FirstSmallestVar / MiddleVal on pcp_domains_read / pcp_domains_size_read,
 *      Branch::distribute (pcp_label) -- on Interval and on IntervalSet domains;
 *   2. pcp_search_run (the C++ driver inside the library);
 *   3. two engines on the two halves of the tree through pcp_consistency_batch.
 * Exit code 0 iff every count is 92 and the contract violations below return PCP_ERR_INVALID. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pcp_b200.h"

#define N 8
#define CHECK(expr)                                                                    \
  do {                                                                                 \
    int _rc = (expr);                                                                  \
    if (_rc != PCP_OK) { fprintf(stderr, "%s -> %d (%s)\n", #expr, _rc, pcp_last_error(e)); return 1; } \
  } while (0)

typedef struct { uint64_t label; int32_t var, val, alt; } branch_t;

static int load_nqueens(pcp_engine* e) {
  int32_t lo[N], hi[N];
  for (int i = 0; i < N; ++i) { lo[i] = 1; hi[i] = N; }
  CHECK(pcp_vars_alloc(e, lo, hi, N, NULL));
  /* example/src/nqueens.rs:36-50: the two diagonals per pair, then join_distinct */
  for (int i = 0; i < N - 1; ++i)
    for (int j = i + 1; j < N; ++j) {
      pcp_operand a[2] = {{i, 0}, {j, j - i}}, b[2] = {{i, 0}, {j, -(j - i)}};
      CHECK(pcp_prop_alloc(e, PCP_X_NEQ_Y, a, 2, NULL));
      CHECK(pcp_prop_alloc(e, PCP_X_NEQ_Y, b, 2, NULL));
    }
  for (int i = 0; i < N - 1; ++i)
    for (int j = i + 1; j < N; ++j) {
      pcp_operand a[2] = {{i, 0}, {j, 0}};
      CHECK(pcp_prop_alloc(e, PCP_X_NEQ_Y, a, 2, NULL));
    }
  return 0;
}

static int post(pcp_engine* e, const branch_t* b) { /* binary_split.rs:46-57 */
  pcp_operand le[2] = {{b->var, 0}, {PCP_VAR_CONSTANT, b->val + 1}}, gt[2] = {{PCP_VAR_CONSTANT, b->val}, {b->var, 0}};
  CHECK(pcp_prop_alloc(e, PCP_X_LESS_Y, b->alt == 0 ? le : gt, 2, NULL));
  return 0;
}

/* status of the node just evaluated -> push its children; returns 1 on a solution */
static int after_node(pcp_engine* e, int32_t status, branch_t* stack, int* sp) {
  if (status == PCP_TRUE) return 1;
  if (status == PCP_FALSE) return 0;
  int32_t lo[N], hi[N];
  uint32_t size[N];
  if (pcp_domains_read(e, 0, N, lo, hi) != PCP_OK || pcp_domains_size_read(e, 0, N, size) != PCP_OK) return -1;
  int var = -1;
  for (int i = 0; i < N; ++i)
    if (size[i] > 1 && (var < 0 || size[i] < size[var])) var = i; /* first_smallest_var.rs:30-39 */
  if (var < 0) return -1;
  uint64_t label;
  if (pcp_label(e, &label) != PCP_OK) return -1;
  branch_t b = {label, var, (lo[var] + hi[var]) / 2, 1};           /* middle_val.rs:25-27 */
  stack[(*sp)++] = b;                                              /* right pushed first: left explored first */
  b.alt = 0;
  stack[(*sp)++] = b;
  return 0;
}

static int host_driven(uint32_t flags, long* solutions, long* nodes) {
  pcp_config cfg = {0, flags, 0, 0};
  pcp_engine* e = NULL;
  if (pcp_engine_create(&cfg, &e) != PCP_OK) return 1;
  if (load_nqueens(e)) return 1;
  static branch_t stack[4096];
  int sp = 0, started = 0;
  *solutions = *nodes = 0;
  while (!started || sp > 0) {
    if (started) {
      branch_t b = stack[--sp];
      CHECK(pcp_restore(e, b.label));                             /* branch.rs:51-55 */
      if (post(e, &b)) return 1;
    }
    started = 1;
    int32_t status;
    pcp_stats st;
    CHECK(pcp_consistency(e, &status, &st));                      /* propagation.rs:49 */
    ++*nodes;
    int r = after_node(e, status, stack, &sp);
    if (r < 0) return 1;
    *solutions += r;
  }
  /* contract violations are error codes, never aborts */
  pcp_operand bad[2] = {{0, 0}, {99, 0}};
  int32_t lo1 = 3, hi1 = 2;
  if (pcp_prop_alloc(e, PCP_X_LESS_Y, bad, 2, NULL) != PCP_ERR_INVALID) return 1;
  if (pcp_vars_alloc(e, &lo1, &hi1, 1, NULL) != PCP_ERR_INVALID) return 1;
  if (pcp_restore(e, 123456) != PCP_ERR_INVALID) return 1;
  pcp_engine_destroy(e);
  return 0;
}

static int library_search(long* solutions) {
  pcp_config cfg = {0, 0, 0, 0};
  pcp_engine* e = NULL;
  if (pcp_engine_create(&cfg, &e) != PCP_OK) return 1;
  if (load_nqueens(e)) return 1;
  pcp_search_config sc;
  memset(&sc, 0, sizeof sc);
  sc.all_solutions = 1;
  pcp_search_result res;
  CHECK(pcp_search_run(e, &sc, &res, NULL, NULL, NULL, NULL, 0));
  *solutions = (long)res.num_solution;
  pcp_engine_destroy(e);
  return res.status == 2 ? 0 : 1;
}

/* two engines, one per child of the root, advanced in lockstep with pcp_consistency_batch */
static int two_halves(long* solutions) {
  pcp_config cfg = {0, 0, 0, 0};
  pcp_engine* es[2] = {NULL, NULL};
  static branch_t stacks[2][4096];
  int sp[2] = {0, 0}, live[2] = {1, 1};
  *solutions = 0;
  for (int k = 0; k < 2; ++k) {
    pcp_engine* e = NULL;
    if (pcp_engine_create(&cfg, &e) != PCP_OK) return 1;
    es[k] = e;
    if (load_nqueens(e)) return 1;
    CHECK(pcp_set_grid_limit(e, 64));
    branch_t root = {0, 0, (1 + N) / 2, k};                       /* the root's split: q0 <= 4 | q0 > 4 */
    if (post(e, &root)) return 1;
  }
  while (live[0] || live[1]) {
    pcp_engine* batch[2];
    int who[2], n = 0;
    for (int k = 0; k < 2; ++k)
      if (live[k]) { batch[n] = es[k]; who[n++] = k; }
    int32_t status[2];
    if (pcp_consistency_batch(batch, n, status, NULL) != PCP_OK) return 1;
    for (int i = 0; i < n; ++i) {
      int k = who[i];
      pcp_engine* e = es[k];
      int r = after_node(e, status[i], stacks[k], &sp[k]);
      if (r < 0) return 1;
      *solutions += r;
      if (sp[k] == 0) { live[k] = 0; continue; }
      branch_t b = stacks[k][--sp[k]];
      CHECK(pcp_restore(e, b.label));
      if (post(e, &b)) return 1;
    }
  }
  pcp_engine_destroy(es[0]);
  pcp_engine_destroy(es[1]);
  return 0;
}

int main(void) {
  long s_iv = 0, n_iv = 0, s_set = 0, n_set = 0, s_lib = 0, s_two = 0;
  if (host_driven(0, &s_iv, &n_iv)) { fprintf(stderr, "host-driven loop (Interval) failed\n"); return 1; }
  if (host_driven(PCP_FLAG_INTERVAL_SET, &s_set, &n_set)) { fprintf(stderr, "host-driven loop (IntervalSet) failed\n"); return 1; }
  if (library_search(&s_lib)) { fprintf(stderr, "pcp_search_run failed\n"); return 1; }
  if (two_halves(&s_two)) { fprintf(stderr, "pcp_consistency_batch loop failed\n"); return 1; }
  printf("nqueens(%d): host-driven Interval %ld solutions / %ld nodes, IntervalSet %ld / %ld, pcp_search_run %ld, two engines %ld\n",
         N, s_iv, n_iv, s_set, n_set, s_lib, s_two);
  return (s_iv == 92 && s_set == 92 && s_lib == 92 && s_two == 92 && n_set < n_iv) ? 0 : 2;
}
