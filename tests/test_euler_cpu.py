"""Eulerian-path serialisation: the multi-threaded C++ host routine behind ggpt_euler_paths against the
property-level oracle (oracle/euler_oracle.py) and, in the build container, against the reference's own functions
(nx_utils.connected_graph2path / shorten_path / get_structure_raw_node2idx_mapping).  Host code only: runs without a GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _random_graph(rng, n, extra, n_comp=1):
    """Bounded-degree trees + ring closures per component (molecule-like), nodes of the components interleaved."""
    perm = rng.permutation(n)
    sizes = np.full(n_comp, n // n_comp)
    sizes[: n % n_comp] += 1
    edges, o = [], 0
    for sz in sizes:
        nodes = perm[o:o + sz]
        o += sz
        for i in range(1, sz):
            edges.append((nodes[i], nodes[rng.integers(max(0, i - 4), i)]))
        for _ in range(extra):
            if sz > 2:
                a, b = rng.choice(sz, 2, replace=False)
                edges.append((nodes[a], nodes[b]))
    return np.asarray(edges, dtype=np.int64).reshape(-1, 2)


def test_walk_properties_on_random_graphs():
    from graphgpt_b200 import euler
    from oracle import euler_oracle
    rng = np.random.default_rng(0)
    graphs = []
    for i in range(300):
        n = int(rng.integers(1, 40))
        e = euler.dedup_edges(_random_graph(rng, n, int(rng.integers(0, 4)), n_comp=int(rng.integers(1, min(n, 4) + 1))), n)
        graphs.append((n, e))
    graphs.append((1, np.zeros((0, 2), np.int32)))                   # single node: empty walk
    graphs.append((2, np.asarray([[0, 1]], np.int32)))               # single edge
    graphs.append((4, np.zeros((0, 2), np.int32)))                   # four isolated nodes: three jump edges
    steps, maps = euler.euler_paths(graphs, seed=7, scope=512)
    for (n, e), st, mp in zip(graphs, steps, maps):
        euler_oracle.check_walk(n, e, st)
        # P6: the re-index equals the reference rule for the start the routine drew (read off the first node)
        first = int(st[0][0]) if len(st) else 0
        ref = euler_oracle.cyclic_map(st, n, int(mp[first]), 512)
        assert all(int(mp[k]) == v for k, v in ref.items())
        assert sorted(ref) == list(range(n)) or len(st) == 0
    # deterministic in (seed, graph index), independent of the thread count; different seeds give different walks
    s1, m1 = euler.euler_paths(graphs, seed=7, scope=512, n_threads=1)
    assert all(np.array_equal(a, b) for a, b in zip(steps, s1)) and all(np.array_equal(a, b) for a, b in zip(maps, m1))
    s2, _ = euler.euler_paths(graphs, seed=8, scope=512)
    assert any(not np.array_equal(a, b) for a, b in zip(steps, s2))


def test_walk_lengths_match_the_pcqm_shaped_statistics():
    """Rows per sample = walk length + 1 (+ <eos>): molecule-like graphs must give the length distribution the synthetic
    workload assumes (mean 23.5 rows) — the eulerisation may not inflate the walks."""
    from graphgpt_b200 import euler
    rng = np.random.default_rng(1)
    graphs = []
    for _ in range(2000):
        n = int(np.clip(np.rint(rng.normal(14.1, 2.5)), 2, 20))
        graphs.append((n, euler.dedup_edges(_random_graph(rng, n, int(rng.integers(0, 4))), n)))
    steps, _ = euler.euler_paths(graphs, seed=3)
    rows = np.asarray([len(s) + 2 for s in steps])
    edges = np.asarray([len(g[1]) for g in graphs])
    assert np.all(np.asarray([len(s) for s in steps]) >= edges)            # every edge at least once
    assert rows.mean() < 1.6 * (edges.mean() + 2), (rows.mean(), edges.mean())


_REF_SCRIPT = r'''
import sys, random
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/baseline")
import ref_loader
ref_loader.load_reference()
import numpy as np, networkx as nx
from src.utils import nx_utils
from graphgpt_b200 import euler
from oracle import euler_oracle

rng = np.random.default_rng(5)
ours, refs = [], []
graphs = []
for i in range(400):
    n = int(rng.integers(3, 30))
    edges = [(j, int(rng.integers(max(0, j - 4), j))) for j in range(1, n)]
    for _ in range(int(rng.integers(0, 4))):
        a, b = rng.choice(n, 2, replace=False)
        edges.append((int(a), int(b)))
    e = euler.dedup_edges(np.asarray(edges), n)
    graphs.append((n, e))
    G = nx.Graph(); G.add_nodes_from(range(n)); G.add_edges_from([tuple(map(int, x)) for x in e])
    random.seed(i)
    path = nx_utils.connected_graph2path(G)                       # the reference's own walk of a connected graph
    steps = np.asarray([(s, t, 0) for s, t in path], dtype=np.int64).reshape(-1, 3)
    eset = {(min(u, v), max(u, v)): k for k, (u, v) in enumerate(e.tolist())}
    steps[:, 2] = [eset[(min(s, t), max(s, t))] for s, t in path]
    euler_oracle.check_walk(n, e, steps)                          # the oracle's properties hold for the REFERENCE
    refs.append(len(path))
    # P6 bit-exact against the reference's mapping function for a forced start index
    start = int(rng.integers(0, 512))
    orig = random.randint
    random.randint = lambda a, b: start
    try:
        ref_map = nx_utils.get_structure_raw_node2idx_mapping(path, 0, 512, 1)
    finally:
        random.randint = orig
    mine = euler_oracle.cyclic_map(steps, n, start, 512)
    assert {k: int(v) for k, v in ref_map.items()} == mine
st, _ = euler.euler_paths(graphs, seed=11)
ours = [len(s) for s in st]
ratio = np.mean(ours) / np.mean(refs)
print(f"mean walk length: native {np.mean(ours):.2f}, reference (networkx eulerize) {np.mean(refs):.2f}, ratio {ratio:.3f}")
assert 0.9 <= ratio <= 1.1, ratio        # greedy shortest-path pairing vs optimal matching: same lengths within 10 %
print("euler reference ok")
'''


@pytest.mark.skipif(not (os.path.isdir("/root/reference/src") or os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "src"))),
                    reason="needs the reference package")
def test_against_the_reference_walk_and_mapping(tmp_path):
    script = tmp_path / "euler_ref.py"
    script.write_text(_REF_SCRIPT)
    p = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "euler reference ok" in p.stdout, (p.stdout + p.stderr)[-3000:]
    print(p.stdout.strip().splitlines()[-2])
