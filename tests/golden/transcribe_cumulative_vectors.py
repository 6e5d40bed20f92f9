#!/usr/bin/env python
"""Transcribes the reference's Cumulative tests -- the only live tests of Store::consistency in
libpcp (src/libpcp/propagators/cumulative.rs:254-319, fixture :151-252) -- into
tests/golden/cumulative_vectors.json.  Hand-transcribed from the Rust source (the reference cannot
be built or imported here); each vector: the task domains, the capacity, whether assigned
domains become Constant views (`constant`), Store::is_subsumed before consistency(), the result
of consistency(), which the fixture also asserts for Store::is_subsumed afterwards."""
import json
import os

U, T, F = 0, 1, -1


def assignment(starts, durations, resources, capacity):
    return {"starts": [[s, s] for s in starts], "durations": [[d, d] for d in durations],
            "resources": [[r, r] for r in resources], "capacity": [capacity, capacity]}


def vec(name, ref, inst, before, after, constant):
    v = dict(inst)
    v.update(name=name, ref=ref, before=before, after=after, constant=constant)
    return v


V = []
# cumulative.rs:254-258 disjunctive_test
V.append(vec("disjunctive", "cumulative.rs:254-258", assignment([0, 0], [0, 0], [1, 1], 1), U, T, False))
# cumulative.rs:260-265 singleton_task
V.append(vec("singleton", "cumulative.rs:260-265", assignment([0], [0], [1], 1), T, T, False))
V.append(vec("singleton-constant", "cumulative.rs:260-265", assignment([0], [0], [1], 1), T, T, True))
# cumulative.rs:267-289 cumulative_assignment_test
V.append(vec("assignment-1-overload", "cumulative.rs:270-272", assignment([0, 1, 4], [3, 4, 2], [1, 2, 2], 3), U, F, False))
V.append(vec("assignment-2-delayed", "cumulative.rs:274-276", assignment([0, 1, 5], [3, 4, 2], [1, 2, 2], 3), U, T, False))
V.append(vec("assignment-3-less-resource", "cumulative.rs:278-280", assignment([0, 1, 4], [3, 4, 2], [1, 2, 1], 3), U, T, False))
V.append(vec("assignment-4-more-capacity", "cumulative.rs:282-284", assignment([0, 1, 4], [3, 4, 2], [1, 2, 2], 4), U, T, False))
V.append(vec("assignment-5-shorter", "cumulative.rs:286-288", assignment([0, 1, 4], [3, 3, 2], [1, 2, 2], 3), U, T, False))
# cumulative.rs:291-319 cumulative_test_constant / cumulative_test_variable
for constant in (True, False):
    tag = "constant" if constant else "variable"
    a = assignment([0, 1, 4], [3, 4, 2], [1, 2, 2], 3)
    a["starts"][0] = [0, 4]
    V.append(vec(f"param-1-{tag}", "cumulative.rs:302-306", a, U, F, constant))
    a = assignment([0, 1, 4], [3, 4, 2], [1, 2, 2], 3)
    a["starts"][1] = [0, 1]
    V.append(vec(f"param-2-{tag}", "cumulative.rs:308-312", a, U, U, constant))
    a = assignment([0, 1, 4], [3, 4, 2], [1, 2, 2], 3)
    a["starts"][2] = [4, 5]
    V.append(vec(f"param-3-{tag}", "cumulative.rs:314-318", a, U, U, constant))

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cumulative_vectors.json")
json.dump({"_comment": "Cumulative fixtures of the reference (SKleene: -1 False, 0 Unknown, 1 True)",
           "reference_commit": "9768dd2613dfc02668ef55ba0c114732b82b356a", "vectors": V}, open(out, "w"), indent=1)
print(len(V), "vectors ->", out)
