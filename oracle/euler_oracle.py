"""Property-level oracle for the Eulerian-path serialisation — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's walk (src/utils/nx_utils.py:388-422) is a random object: networkx `eulerize` (blossom matching) plus
Python's `random` stream decide it, so no native routine can reproduce it bit for bit.  What the rest of the pipeline
relies on are the properties below, restated here in plain Python and checked against BOTH the reference's own
`connected_graph2path` / `shorten_path` (tests, build container) and the C++ routine:
  P1  every step (src, tgt) is an edge of the graph, or a jump edge joining the end of one component's walk to the start
      of the next component's walk (nx_utils.py:400-408);
  P2  consecutive steps are chained: tgt_i == src_{i+1};
  P3  every edge of the graph is traversed at least once (in either direction);
  P4  inside a component the walk stops at the step that covers the component's last unvisited edge (shorten_path,
      :331-348): its final step is the first traversal of its edge;
  P5  every component is visited in one contiguous block; single-node components contribute no step of their own;
  P6  cyclic re-index (get_structure_raw_node2idx_mapping, :234-260, mapping_type 1): nodes numbered (start + k) % scope
      in order of first appearance along [src_0, src_1, ..., src_last, tgt_last].
"""
import numpy as np


def components(n_nodes, edges):
    parent = list(range(n_nodes))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for u, v in edges:
        parent[find(int(u))] = find(int(v))
    return [find(i) for i in range(n_nodes)]


def check_walk(n_nodes, edges, steps):
    """edges int [E,2] (simple, undirected), steps int [P,3] = (src, tgt, edge index | -1).  Raises AssertionError."""
    edges = np.asarray(edges).reshape(-1, 2)
    eset = {(min(int(u), int(v)), max(int(u), int(v))): i for i, (u, v) in enumerate(edges)}
    comp = components(n_nodes, edges)
    n_comp = len(set(comp))
    if len(steps) == 0:
        assert len(edges) == 0 and n_comp == 1, "empty walk only for a single component without edges"
        return
    seen_edges, blocks = set(), []
    for i, (s, t, e) in enumerate(steps):
        s, t, e = int(s), int(t), int(e)
        if i > 0:
            assert int(steps[i - 1][1]) == s, f"P2 violated at step {i}"
        key = (min(s, t), max(s, t))
        if e >= 0:
            assert key in eset and eset[key] == e, f"P1: step {i} ({s},{t}) is not edge {e}"
            assert comp[s] == comp[t]
            first = key not in seen_edges
            seen_edges.add(key)
            if not blocks or blocks[-1][0] != comp[s]:
                blocks.append([comp[s], i, i, first])
            blocks[-1][2], blocks[-1][3] = i, first
        else:
            assert comp[s] != comp[t], f"P1: jump edge at step {i} inside one component"
    assert len(seen_edges) == len(eset), f"P3: {len(eset) - len(seen_edges)} edges never traversed"
    visited_comps = [b[0] for b in blocks]
    assert len(set(visited_comps)) == len(visited_comps), "P5: a component is visited in two separate blocks"
    for c, lo, hi, last_is_first in blocks:
        assert last_is_first, f"P4: the walk of component {c} does not stop at its last newly covered edge"
    n_jumps = int(sum(1 for s in steps if int(s[2]) < 0))
    assert n_jumps == n_comp - 1, f"P5: {n_jumps} jump edges for {n_comp} components"


def cyclic_map(steps, n_nodes, start, scope):
    """P6, the reference's get_structure_raw_node2idx_mapping for mapping_type 1 given `start` (its random.randint draw)."""
    if len(steps) == 0:
        return {0: start % scope}
    seq = [int(s[0]) for s in steps] + [int(steps[-1][1])]
    uniq = list(dict.fromkeys(seq))
    return {old: (start + k) % scope for k, old in enumerate(uniq)}
