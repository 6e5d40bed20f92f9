"""Transcription of the known-answer vectors held by libpcp's own unit tests for the
propagation hot path into tests/golden/reference_vectors.json.

The reference is Rust and cannot be executed in this image, so these are hand
transcriptions of the literal test tables (file:line given per group), not generated
outputs.  Run `python tests/golden/transcribe_reference_vectors.py` to regenerate the
JSON; the JSON is what the tests read (this script is never imported by tests).

Vector format (fixture `test_propagation`, src/libpcp/propagators/mod.rs:110-129):
  domains  initial Interval<i32> domains, one VStoreFD variable each, in alloc order
  kind/ops the propagator as a device descriptor: (var, off) operands, var=-1 constant
  before   is_subsumed before propagate      (-1 False, 0 Unknown, 1 True)
  ok       propagate() return value
  delta    exact ordered delta [(var, event)] (0 Assignment, 1 Bound, 2 Inner), checked
           only when ok (mod.rs:124-126)
  after    is_subsumed after propagate
"""
import json
import os

U, T, F = 0, 1, -1
A, B = 0, 1  # Assignment, Bound

d0_10, d10_20, d10_11, d5_15, d11_20, d1_1 = [0, 10], [10, 20], [10, 11], [5, 15], [11, 20], [1, 1]
d1_10, d5_10, d6_10, d2_2 = [1, 10], [5, 10], [6, 10], [2, 2]
d11_12, d0_6, d0_5, d0_1, d12_12 = [11, 12], [0, 6], [0, 5], [0, 1], [12, 12]
zero, one, two, ten = [0, 0], [1, 1], [2, 2], [10, 10]
d0_2, d0_3 = [0, 2], [0, 3]

X, Y, Z = [0, 0], [1, 0], [2, 0]

vectors = []


def add(group, ref, num, kind, domains, ops, before, after, delta, ok):
    vectors.append({"name": f"{group}/{num}", "ref": ref, "kind": kind, "domains": domains, "ops": ops,
                    "before": before, "after": after, "delta": [list(d) for d in delta], "ok": ok})


# --- src/libpcp/propagators/cmp/x_less_y.rs:126-167
ref = "propagators/cmp/x_less_y.rs:126-167"
for num, x, y, b, a, d, ok in [
    (1, d0_10, d0_10, U, U, [(0, B), (1, B)], True),
    (2, d0_10, d10_20, U, U, [], True),
    (3, d10_11, d10_11, U, T, [(0, A), (1, A)], True),
    (4, d5_15, d10_20, U, U, [], True),
    (5, d5_15, d0_10, U, U, [(0, B), (1, B)], True),
    (6, d0_10, d11_20, T, T, [], True),
    (7, d11_20, d0_10, F, F, [], False),
    (8, d1_1, d0_10, U, T, [(1, B)], True),
]:
    add("x_less_y", ref, num, "XLessY", [x, y], [X, Y], b, a, d, ok)

# --- src/libpcp/term/addition.rs:128-170   XLessY(x, Addition(y, c))
ref = "term/addition.rs:128-170"
for num, x, y, c, b, a, d, ok in [
    (1, d0_10, d5_15, -5, U, U, [(0, B), (1, B)], True),
    (2, d0_10, d0_10, 10, U, U, [], True),
    (3, d5_15, d5_15, 5, U, U, [], True),
    (4, d5_15, d10_20, -10, U, U, [(0, B), (1, B)], True),
    (5, d0_10, d0_10, 11, T, T, [], True),
    (6, d0_10, d0_10, -11, F, F, [], False),
    (7, d1_1, d5_15, -5, U, T, [(1, B)], True),
]:
    add("x_less_y_plus_c", ref, num, "XLessY", [x, y], [X, [1, c]], b, a, d, ok)

# --- src/libpcp/term/constant.rs:86-103   XLessY(x, Constant(5)) and read-back [0,4]
add("x_less_constant", "term/constant.rs:86-103", 1, "XLessY", [d0_10], [X, [-1, 5]], U, T, [(0, B)], True)
vectors[-1]["domains_after"] = [[0, 4]]

# --- src/libpcp/term/constant.rs:105-186   unary propagators over Constant
ref = "term/constant.rs:105-186"
# lowering of the derived constructors (propagators/cmp/mod.rs:40-60):
#   XLessY::new(x, c)        -> XLessY(x, c)
#   x_leq_y(x, c)            -> XLessY(x, Addition(c, 1))      = (x, const c+1)
#   x_greater_y(x, c)        -> XLessY(c, x)
#   x_geq_y(x, c)            -> XLessY(c, Addition(x, 1))
dom0_0 = [0, 0]
for num, x, c, ctor, b, a, d, ok in [
    (1, d0_10, 0, "less", F, F, [], False),
    (2, d0_10, 11, "less", T, T, [], True),
    (3, d0_10, 10, "less", U, T, [(0, B)], True),
    (4, d0_10, -1, "leq", F, F, [], False),
    (5, d0_10, 10, "leq", T, T, [], True),
    (6, d0_10, 9, "leq", U, T, [(0, B)], True),
    (7, d0_10, 10, "greater", F, F, [], False),
    (8, d0_10, -1, "greater", T, T, [], True),
    (9, d0_10, 0, "greater", U, T, [(0, B)], True),
    (10, d0_10, 11, "geq", F, F, [], False),
    (11, d0_10, 0, "geq", T, T, [], True),
    (12, d0_10, 1, "geq", U, T, [(0, B)], True),
    (13, dom0_0, 0, "neq", F, F, [], False),
    (14, d0_10, 5, "neq", U, U, [], True),
    (15, d0_10, 0, "neq", U, T, [(0, B)], True),
    (16, d0_10, 10, "neq", U, T, [(0, B)], True),
]:
    kind, ops = {
        "less": ("XLessY", [X, [-1, c]]),
        "leq": ("XLessY", [X, [-1, c + 1]]),
        "greater": ("XLessY", [[-1, c], X]),
        "geq": ("XLessY", [[-1, c], [0, 1]]),
        "neq": ("XNeqY", [X, [-1, c]]),
    }[ctor]
    add("unary_" + ctor, ref, num, kind, [x], ops, b, a, d, ok)

# --- src/libpcp/propagators/cmp/x_greater_y_plus_z.rs:137-182
ref = "propagators/cmp/x_greater_y_plus_z.rs:137-182"
for num, x, y, z, b, a, d, ok in [
    (1, d0_10, d0_10, d0_10, U, U, [(0, B), (1, B), (2, B)], True),
    (2, d10_11, d5_15, d5_15, U, T, [(0, A), (1, A), (2, A)], True),
    (3, d10_20, d1_1, d1_1, T, T, [], True),
    (4, d1_1, d1_1, d1_1, F, F, [], False),
    (5, d2_2, d1_1, d1_1, F, F, [], False),
    (6, d6_10, d5_10, d1_10, U, U, [(0, B), (1, B), (2, B)], True),
]:
    add("x_greater_y_plus_z", ref, num, "XGreaterYPlusZ", [x, y, z], [X, Y, Z], b, a, d, ok)

# --- src/libpcp/propagators/cmp/x_less_y_plus_z.rs:137-172
ref = "propagators/cmp/x_less_y_plus_z.rs:137-172"
for num, x, y, z, b, a, d, ok in [
    (1, d0_10, d0_10, d0_10, U, U, [], True),
    (2, d11_12, d0_6, d0_6, U, T, [(0, A), (1, A), (2, A)], True),
    (3, d10_20, d1_1, d1_1, F, F, [], False),
    (4, d2_2, d1_1, d1_1, F, F, [], False),
    (5, d1_1, d2_2, d2_2, T, T, [], True),
    (6, d0_6, d0_5, d0_1, U, U, [(0, B)], True),
]:
    add("x_less_y_plus_z", ref, num, "XLessYPlusZ", [x, y, z], [X, Y, Z], b, a, d, ok)

# --- src/libpcp/propagators/cmp/x_eq_y_plus_z.rs:115-141
ref = "propagators/cmp/x_eq_y_plus_z.rs:115-141"
for num, x, y, z, b, a, d, ok in [
    (1, d0_10, d0_10, d0_10, U, U, [], True),
    (2, d12_12, d0_6, d0_6, U, T, [(1, A), (2, A)], True),
    (3, d10_20, d1_1, d1_1, F, F, [], False),
    (4, d2_2, d1_1, d1_1, T, T, [], True),
    (5, d1_1, d2_2, d2_2, F, F, [], False),
    (6, d0_6, d0_5, d0_1, U, U, [], True),
]:
    add("x_eq_y_plus_z", ref, num, "XEqYPlusZ", [x, y, z], [X, Y, Z], b, a, d, ok)

# --- src/libpcp/propagators/cmp/x_eq_y_mul_z.rs:124-147
ref = "propagators/cmp/x_eq_y_mul_z.rs:124-147"
d1_2 = [1, 2]
for num, x, y, z, b, a, d, ok in [
    (1, d0_10, d0_10, d0_10, U, U, [], True),
    (2, d10_11, d5_15, d5_15, F, F, [], False),
    (3, d10_20, d1_1, d1_1, F, F, [], False),
    (4, d1_1, d1_1, d1_1, T, T, [], True),
    (5, d1_2, d1_1, d1_1, U, T, [(0, A)], True),
]:
    add("x_eq_y_mul_z", ref, num, "XEqYMulZ", [x, y, z], [X, Y, Z], b, a, d, ok)

# --- src/libpcp/propagators/cmp/x_neq_y.rs:114-133
ref = "propagators/cmp/x_neq_y.rs:114-133"
for num, x, y, b, a, d, ok in [
    (1, d0_10, d0_10, U, U, [], True),
    (2, d0_10, d10_20, U, U, [], True),
    (3, d5_15, d10_20, U, U, [], True),
    (4, d0_10, d11_20, T, T, [], True),
    (5, one, d0_10, U, U, [], True),
    (6, zero, d0_10, U, T, [(1, B)], True),
    (7, ten, d0_10, U, T, [(1, B)], True),
    (8, one, one, F, F, [], False),
    (9, zero, one, T, T, [], True),
]:
    add("x_neq_y", ref, num, "XNeqY", [x, y], [X, Y], b, a, d, ok)

# --- src/libpcp/propagators/cmp/x_eq_y.rs:125-172
ref = "propagators/cmp/x_eq_y.rs:125-172"
for num, x, y, b, a, d, ok in [
    (1, d0_10, d0_10, U, U, [], True),
    (2, d0_10, d10_20, U, T, [(0, A), (1, A)], True),
    (3, d5_15, d10_20, U, U, [(0, B), (1, B)], True),
    (4, d5_15, d0_10, U, U, [(0, B), (1, B)], True),
    (5, d0_10, d11_20, F, F, [], False),
    (6, d11_20, d0_10, F, F, [], False),
    (7, d1_1, d0_10, U, T, [(1, A)], True),
]:
    add("x_eq_y", ref, num, "XEqY", [x, y], [X, Y], b, a, d, ok)

# --- src/libpcp/propagators/distinct.rs:136-173
ref = "propagators/distinct.rs:136-173"
for num, doms, b, a, d, ok in [
    (1, [zero, one, two], T, T, [], True),
    (2, [zero, zero, two], F, F, [], False),
    (3, [zero, one, d0_3], U, T, [(2, B)], True),
    (4, [zero, one, d0_2], U, T, [(2, A)], True),
    (5, [zero, one, d0_1], U, F, [], False),
    (6, [zero, d0_3, d0_3], U, U, [(1, B), (2, B)], True),
    (7, [d0_3], T, T, [], True),
]:
    add("distinct", ref, num, "Distinct", doms, [[i, 0] for i in range(len(doms))], b, a, d, ok)

# --- src/libpcp/propagators/all_equal.rs:112-151
ref = "propagators/all_equal.rs:112-151"
for num, doms, b, a, d, ok in [
    (1, [zero, one, two], F, F, [], False),
    (2, [zero, zero, two], F, F, [], False),
    (3, [zero, zero, zero], T, T, [], True),
    (4, [zero, d0_3, d0_3], U, T, [(1, A), (2, A)], True),
    (5, [d0_1, d0_3, d0_3], U, U, [(1, B), (2, B)], True),
    (6, [zero, one, d0_2], F, F, [], False),
    (7, [d0_1, one, d0_1], U, T, [(0, A), (2, A)], True),
    (8, [d0_3], T, T, [], True),
    (9, [one], T, T, [], True),
]:
    add("all_equal", ref, num, "AllEqual", doms, [[i, 0] for i in range(len(doms))], b, a, d, ok)

# --- variable store semantics: src/libpcp/variable/store.rs:422-526 ---------------------
store_updates = []
for ref, num, source, target, events, ok in [
    ("variable/store.rs:422-436", 1, d0_10, d0_10, [], True),
    ("variable/store.rs:422-436", 2, d0_10, [1, 0], [], False),
    ("variable/store.rs:422-436", 3, d0_10, [0, 0], [A], True),
    ("variable/store.rs:422-436", 4, d0_10, [5, 10], [B], True),
    ("variable/store.rs:422-436", 5, d0_10, [0, 5], [B], True),
    ("variable/store.rs:422-436", 6, d0_10, [1, 9], [B], True),
    # var_shrink_bound (store.rs:456-468): shrink_left / shrink_right results
    ("variable/store.rs:456-468", 1, d0_10, [0, 10], [], True),
    ("variable/store.rs:456-468", 2, d0_10, [10, 10], [A], True),
    ("variable/store.rs:456-468", 3, d0_10, [1, 10], [B], True),
    ("variable/store.rs:456-468", 4, d0_10, [11, 10], [], False),
    ("variable/store.rs:456-468", 5, d0_10, [0, 10], [], True),
    ("variable/store.rs:456-468", 6, d0_10, [0, 0], [A], True),
    ("variable/store.rs:456-468", 7, d0_10, [0, 1], [B], True),
    ("variable/store.rs:456-468", 8, d0_10, [0, -1], [], False),
]:
    store_updates.append({"ref": ref, "num": num, "source": source, "target": target, "events": events, "ok": ok})

# var_intersection_test (store.rs:509-526): both vars updated with x ∩ y  == XEqY.propagate
store_intersections = [
    {"ref": "variable/store.rs:509-526", "x": d0_10, "y": d10_20, "target": [10, 10], "delta": [[0, A], [1, A]], "ok": True},
    {"ref": "variable/store.rs:509-526", "x": d0_10, "y": [1, 9], "target": [1, 9], "delta": [[0, B]], "ok": True},
    {"ref": "variable/store.rs:509-526", "x": [1, 9], "y": d0_10, "target": [1, 9], "delta": [[1, B]], "ok": True},
    {"ref": "variable/store.rs:509-526", "x": d0_10, "y": d11_20, "target": [1, 0], "delta": [], "ok": False},
]

# --- search goldens -----------------------------------------------------------------------
search = {
    "nqueens_all_solutions": {"ref": "search/engine/all_solution.rs:67-74", "counts": [1, 0, 0, 2, 10, 4, 40, 92, 352]},
    "nqueens_one_solution": {"ref": "search/engine/one_solution.rs:120-128",
                             "status": {"1": 1, "2": -1, "3": -1, "4": 1, "5": 1, "6": 1, "7": 1, "8": 1, "9": 1,
                                        "10": 1, "11": 1}},
    "stop_node": {"ref": "search/stop_node.rs:82-104", "n": 6, "limit": 10, "num_nodes": 10, "status": 2},
    "branch_and_bound": {"ref": "search/branch_and_bound.rs:112-138", "maximize": 9, "minimize": 0},
    "binary_split": {"ref": "search/branching/binary_split.rs:110-134", "root": [[1, 10], [2, 4], [1, 2]],
                     "children": {"0": [[1, 5], [6, 10]], "1": [[2, 3], [4, 4]], "2": [[1, 1], [2, 2]]}},
    "first_smallest_var": {"ref": "search/branching/first_smallest_var.rs:63-72",
                           "cases": [{"vars": [[1, 10], [2, 4], [1, 1]], "expect": 1},
                                     {"vars": [[1, 10], [2, 4], [2, 4]], "expect": 1},
                                     {"vars": [[1, 1], [1, 1], [1, 10], [1, 1], [2, 4], [1, 1], [1, 1]], "expect": 4}]},
    # dead tests, still informative: propagation/store.rs:362-392
    "chained_lt": {"ref": "propagation/store.rs:362-384", "status": {"1": 1, "2": 0, "5": 0, "9": 0, "10": 1, "11": -1}},
    "nqueens_root": {"ref": "propagation/store.rs:386-392", "status": {"1": 1, "2": 0, "3": 0, "4": 0}},
}

out = {
    "_comment": "Hand-transcribed from libpcp's unit tests; see transcribe_reference_vectors.py",
    "reference_commit": "9768dd2613dfc02668ef55ba0c114732b82b356a",
    "propagators": vectors,
    "store_updates": store_updates,
    "store_intersections": store_intersections,
    "search": search,
}

if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {path}: {len(vectors)} propagator vectors")
