"""Eulerian-path serialisation of graph batches through the C ABI.

  euler_paths(graphs, seed, scope)      host, multi-threaded C++ (ggpt_euler_paths): the walk + cyclic node re-index of
                                        nx_utils.graph2path_v2 / get_structure_raw_node2idx_mapping for a whole batch
  stack_rows(...)                       device gather (ggpt_stack_path_rows): the stacked token rows of
                                        tokenizer.stack_node_edge_graph_attr_to_node, written straight into the row pool
                                        that graphgpt_b200.packing / ggpt_pack_sequences consume
Together with ggpt_pack_sequences and ggpt_smtp_mask_2d this takes graphs -> packed, masked training batches without the
per-sample networkx / Python-dict work of the reference's DataLoader workers.
"""
import ctypes

import numpy as np

from .lib import lib


def dedup_edges(edges, n_nodes):
    """Undirected simple edge list in first-occurrence order (what the reference's to_networkx(...).to_undirected()
    leaves): self loops dropped, (u, v) and (v, u) merged.  Returns int32 [E, 2] with u < v."""
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    e = e[e[:, 0] != e[:, 1]]
    lo, hi = np.minimum(e[:, 0], e[:, 1]), np.maximum(e[:, 0], e[:, 1])
    key = lo * n_nodes + hi
    _, first = np.unique(key, return_index=True)
    first.sort()
    return np.stack([lo[first], hi[first]], axis=1).astype(np.int32)


def euler_paths(graphs, seed=0, scope=512, n_threads=0):
    """graphs: list of (n_nodes, edges int [E,2]) with edges already de-duplicated (dedup_edges).
    Returns (steps, node_maps): per graph an int32 [P,3] array of (src, tgt, edge index | -1 for a jump edge) and an
    int32 [n_nodes] array node -> re-indexed id in [0, scope)."""
    G = len(graphs)
    node_count = np.asarray([g[0] for g in graphs], dtype=np.int32)
    edge_off = np.zeros(G + 1, dtype=np.int64)
    edge_off[1:] = np.cumsum([len(g[1]) for g in graphs])
    edges = (np.concatenate([np.asarray(g[1], dtype=np.int32).reshape(-1, 2) for g in graphs], axis=0)
             if edge_off[-1] > 0 else np.zeros((1, 2), np.int32))
    edges = np.ascontiguousarray(edges)
    step_off = np.zeros(G + 1, dtype=np.int64)
    node_off = np.zeros(G + 1, dtype=np.int64)
    node_map = np.zeros(int(node_count.sum()), dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    total = lib.ggpt_euler_paths(G, p(node_count), p(edge_off), p(edges), int(seed), int(scope), int(n_threads), p(step_off),
                                 None, 0, p(node_off), p(node_map))
    if total < 0:
        raise RuntimeError(f"ggpt_euler_paths failed: {lib.last_error()}")
    steps = np.zeros((max(int(total), 1), 3), dtype=np.int32)
    # the routine is deterministic in (seed, graph): the second call reproduces the walk and copies it out
    total2 = lib.ggpt_euler_paths(G, p(node_count), p(edge_off), p(edges), int(seed), int(scope), int(n_threads), p(step_off),
                                  p(steps), int(total), p(node_off), p(node_map))
    assert total2 == total
    out_steps = [steps[step_off[g]:step_off[g + 1]] for g in range(G)]
    out_maps = [node_map[node_off[g]:node_off[g + 1]] for g in range(G)]
    return out_steps, out_maps


def path_rows(steps, n_nodes_offset=0, edge_offset=0):
    """(row_node, row_edge) of one graph's token rows: the first row is the walk's start node with no edge, then one row
    per step (target node, traversed edge) — tokenizer.py:1207-1263.  Offsets shift into batch-global tables."""
    if len(steps) == 0:
        return np.asarray([n_nodes_offset], np.int32), np.asarray([-1], np.int32)
    node = np.concatenate([steps[:1, 0], steps[:, 1]]).astype(np.int32) + n_nodes_offset
    edge = np.concatenate([[-1], np.where(steps[:, 2] >= 0, steps[:, 2] + edge_offset, -1)]).astype(np.int32)
    return node, edge


def stack_rows(row_node, row_edge, node_map, node_attr, edge_attr, default_edge, node_base):
    """Device gather.  row_node / row_edge int32 [R], node_map int32 [n_nodes], node_attr int64 [n_nodes, An], edge_attr
    int64 [n_edges, Ae], default_edge int64 [Ae] -> rows int64 [R, 1 + An + Ae] (all CUDA tensors)."""
    import torch
    R = row_node.numel()
    An = node_attr.shape[1] if node_attr is not None else 0
    Ae = edge_attr.shape[1] if edge_attr is not None else 0
    rows = torch.empty((R, 1 + An + Ae), device=row_node.device, dtype=torch.int64)
    ptr = lambda t: 0 if t is None else t.data_ptr()
    lib.ggpt_stack_path_rows(row_node.data_ptr(), row_edge.data_ptr(), node_map.data_ptr(), ptr(node_attr), An, ptr(edge_attr),
                             Ae, ptr(default_edge), int(node_base), rows.data_ptr(), R,
                             torch.cuda.current_stream().cuda_stream)
    return rows
