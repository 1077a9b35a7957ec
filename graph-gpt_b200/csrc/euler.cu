// Eulerian-path serialisation of a batch of graphs — HOST routine (multi-threaded C++), plus the device
// gather that stacks node / edge attribute tokens onto the path positions.
//
// Reference (per sample, networkx + Python dicts inside DataLoader workers):
//   graph2path_v2           src/utils/nx_utils.py:388-410   components -> per-component path -> jump edges between them
//   connected_graph2path    :413-422                        eulerize if needed, random start node, Euler circuit, shorten
//   shorten_path            :331-348                        cut right after the step that covers the last unvisited edge
//   get_structure_raw_node2idx_mapping (cyclic) :234-260    node ids in order of first appearance, (start + k) % scope
//   stack_node_edge_graph_attr_to_node  src/data/tokenizer.py:1196-1266   one row per path position: node-id token, the
//                                                           node's attribute tokens, the traversed edge's attribute tokens
// The reference's walk depends on networkx's blossom matching inside `eulerize` and on Python's `random` stream, so it
// is not reproducible bit for bit; this routine guarantees the same PROPERTIES (tests/test_euler_cpu.py, oracle/
// euler_oracle.py): every edge of the graph is traversed, consecutive steps are adjacent, only existing edges are
// duplicated (odd-degree nodes are paired greedily by BFS distance and joined along shortest paths, as eulerize does with
// an optimal matching), the walk stops at the step that covers the last edge, components are visited in random order and
// joined by jump edges, and — given the walk — the cyclic re-index is identical to the reference's.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <queue>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "../../include/ggpt_b200.h"

namespace ggpt {

struct Rng {   // splitmix64: one independent stream per (seed, graph)
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  int below(int n) { return static_cast<int>(next() % static_cast<uint64_t>(n)); }
  template <class T>
  void shuffle(std::vector<T>& v) {
    for (int i = static_cast<int>(v.size()) - 1; i > 0; --i) std::swap(v[i], v[below(i + 1)]);
  }
};

struct Half {   // one direction of an undirected (multi-)edge
  int to;
  int eid;      // id of the undirected edge instance (duplicates get fresh ids)
  int orig;     // index of the original edge it copies
};

// Euler walk of one connected component (node list `comp`, adjacency over the whole graph restricted to it).
// Appends (src, tgt, original edge id) steps to `out`.
static void component_walk(const std::vector<int>& comp, const std::vector<std::pair<int, int>>& edges,
                           const std::vector<std::vector<std::pair<int, int>>>& adj0, Rng& rng,
                           std::vector<int>& scratch_dist, std::vector<int>& scratch_par, std::vector<int>& scratch_pe,
                           std::vector<int>& out) {
  if (comp.size() == 1) return;
  // multigraph adjacency of this component: original edges first
  std::unordered_map<int, std::vector<Half>> adj;
  int next_eid = 0;
  std::vector<int> comp_edges;
  for (int u : comp)
    for (auto [v, e] : adj0[u])
      if (u < v || (u == v)) comp_edges.push_back(e);
  for (int e : comp_edges) {
    const int u = edges[e].first, v = edges[e].second;
    adj[u].push_back({v, next_eid, e});
    adj[v].push_back({u, next_eid, e});
    ++next_eid;
  }
  // ---- eulerize: pair odd-degree nodes (greedy nearest by BFS distance), duplicate the edges of a shortest path
  std::vector<int> odd;
  for (int u : comp)
    if (adj[u].size() & 1) odd.push_back(u);
  rng.shuffle(odd);
  std::vector<char> matched(odd.size(), 0);
  for (size_t i = 0; i < odd.size(); ++i) {
    if (matched[i]) continue;
    const int src = odd[i];
    // BFS over ORIGINAL edges from src until the nearest unmatched odd node
    for (int u : comp) scratch_dist[u] = -1;
    std::queue<int> q;
    q.push(src);
    scratch_dist[src] = 0;
    int found = -1;
    size_t found_j = 0;
    while (!q.empty() && found < 0) {
      const int u = q.front();
      q.pop();
      for (auto [v, e] : adj0[u]) {
        if (scratch_dist[v] >= 0) continue;
        scratch_dist[v] = scratch_dist[u] + 1;
        scratch_par[v] = u;
        scratch_pe[v] = e;
        q.push(v);
      }
      // nearest unmatched odd node among the nodes labelled so far (checked per BFS layer head)
      for (size_t j = i + 1; j < odd.size(); ++j)
        if (!matched[j] && scratch_dist[odd[j]] >= 0 && (found < 0 || scratch_dist[odd[j]] < scratch_dist[found])) {
          found = odd[j];
          found_j = j;
        }
    }
    if (found < 0) continue;   // cannot happen in a connected component (odd nodes come in pairs)
    matched[i] = matched[found_j] = 1;
    for (int v = found; v != src; v = scratch_par[v]) {
      const int u = scratch_par[v], e = scratch_pe[v];
      adj[u].push_back({v, next_eid, e});
      adj[v].push_back({u, next_eid, e});
      ++next_eid;
    }
  }
  // ---- Hierholzer from a random start node, neighbours in random order: vertices in finishing order form an Euler
  //      circuit of the (now even-degree) multigraph; consecutive vertices are joined by one of its edges
  for (auto& kv : adj) rng.shuffle(kv.second);
  std::vector<char> used(next_eid, 0);
  std::unordered_map<int, size_t> pos;
  const int start = comp[rng.below(static_cast<int>(comp.size()))];
  std::vector<int> stack{start};
  std::vector<int> circuit;
  while (!stack.empty()) {
    const int u = stack.back();
    auto& nb = adj[u];
    size_t& p = pos[u];
    while (p < nb.size() && used[nb[p].eid]) ++p;
    if (p == nb.size()) {
      circuit.push_back(u);
      stack.pop_back();
    } else {
      used[nb[p].eid] = 1;
      stack.push_back(nb[p].to);
    }
  }
  // ---- shorten: stop right after the step that covers the last not-yet-visited original edge
  std::vector<char> seen(edges.size(), 0);
  size_t n_unique = 0;
  {
    std::vector<int> ce = comp_edges;
    std::sort(ce.begin(), ce.end());
    n_unique = std::unique(ce.begin(), ce.end()) - ce.begin();
  }
  size_t covered = 0;
  for (size_t i = 1; i < circuit.size() && covered < n_unique; ++i) {
    const int a = circuit[i - 1], b = circuit[i];
    // the original edge behind this step: the graph is simple, so every copy joining a and b has the same original id
    int orig = -1;
    for (const Half& hf : adj[a])
      if (hf.to == b) {
        orig = hf.orig;
        break;
      }
    out.push_back(a);
    out.push_back(b);
    out.push_back(orig);
    if (orig >= 0 && !seen[orig]) {
      seen[orig] = 1;
      ++covered;
    }
  }
}

// One graph: edges are undirected pairs of local node ids in [0, n_nodes).
static void graph_walk(int n_nodes, const int* edge_ptr, int n_edges, uint64_t seed, int scope, std::vector<int>& steps,
                       std::vector<int>& node_map) {
  Rng rng(seed);
  std::vector<std::pair<int, int>> edges;
  edges.reserve(n_edges);
  std::vector<std::vector<std::pair<int, int>>> adj0(n_nodes);
  {
    std::unordered_map<long long, int> dedup;   // the reference builds a simple undirected graph (to_undirected)
    for (int e = 0; e < n_edges; ++e) {
      int u = edge_ptr[2 * e], v = edge_ptr[2 * e + 1];
      if (u == v || u < 0 || v < 0 || u >= n_nodes || v >= n_nodes) continue;
      if (u > v) std::swap(u, v);
      const long long key = static_cast<long long>(u) * n_nodes + v;
      if (dedup.count(key)) continue;
      const int id = static_cast<int>(edges.size());
      dedup[key] = id;
      edges.push_back({u, v});
      adj0[u].push_back({v, id});
      adj0[v].push_back({u, id});
    }
  }
  // connected components
  std::vector<int> comp_of(n_nodes, -1);
  std::vector<std::vector<int>> comps;
  for (int s = 0; s < n_nodes; ++s) {
    if (comp_of[s] >= 0) continue;
    const int c = static_cast<int>(comps.size());
    comps.emplace_back();
    std::vector<int> st{s};
    comp_of[s] = c;
    while (!st.empty()) {
      const int u = st.back();
      st.pop_back();
      comps[c].push_back(u);
      for (auto [v, e] : adj0[u])
        if (comp_of[v] < 0) {
          comp_of[v] = c;
          st.push_back(v);
        }
    }
    std::sort(comps[c].begin(), comps[c].end());
  }
  rng.shuffle(comps);
  std::vector<int> dist(n_nodes), par(n_nodes), pe(n_nodes);
  steps.clear();
  int prev_connect = -1;
  for (size_t ci = 0; ci < comps.size(); ++ci) {
    std::vector<int> sub;
    component_walk(comps[ci], edges, adj0, rng, dist, par, pe, sub);
    const int first = sub.empty() ? comps[ci][0] : sub[0];
    if (ci > 0) {   // jump edge between components (nx_utils.py:400-408), no original edge behind it
      steps.push_back(prev_connect);
      steps.push_back(first);
      steps.push_back(-1);
    }
    steps.insert(steps.end(), sub.begin(), sub.end());
    prev_connect = sub.empty() ? comps[ci][0] : sub[sub.size() - 2];
  }
  // cyclic re-index (nx_utils.py:234-260, mapping_type 1): nodes in order of first appearance along the walk
  node_map.assign(n_nodes, -1);
  const int start_idx = scope > 0 ? rng.below(scope) : 0;
  int k = 0;
  auto visit = [&](int u) {
    if (node_map[u] < 0) node_map[u] = scope > 0 ? (start_idx + k++) % scope : k++;
  };
  if (steps.empty()) {
    visit(0);
  } else {
    for (size_t i = 0; i < steps.size(); i += 3) visit(steps[i]);
    visit(steps[steps.size() - 2]);
  }
}

// rows[r, :] = [ node_base + node_map[node_r] | node_attr[node_r, :] | edge_attr[edge_r, :] or default_edge[:] ]
__global__ void stack_path_rows_kernel(const int* __restrict__ row_node, const int* __restrict__ row_edge,
                                       const int* __restrict__ node_map, const long long* __restrict__ node_attr, int An,
                                       const long long* __restrict__ edge_attr, int Ae,
                                       const long long* __restrict__ default_edge, long long node_base,
                                       long long* __restrict__ rows, long long R) {
  const int F = 1 + An + Ae;
  const long long total = R * F;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / F;
    const int f = static_cast<int>(i % F);
    const int node = row_node[r];
    long long v;
    if (f == 0) v = node_base + node_map[node];
    else if (f <= An) v = node_attr[static_cast<long long>(node) * An + (f - 1)];
    else {
      const int e = row_edge[r];
      v = e >= 0 ? edge_attr[static_cast<long long>(e) * Ae + (f - 1 - An)] : default_edge[f - 1 - An];
    }
    rows[i] = v;
  }
}

}  // namespace ggpt

using namespace ggpt;

extern "C" {

long long ggpt_euler_paths(int n_graphs, const int* node_count, const long long* edge_off, const int* edges,
                           unsigned long long seed, int scope, int n_threads, long long* step_off, int* steps,
                           long long steps_cap, long long* node_off, int* node_map) {
  if (n_graphs <= 0 || !node_count || !edge_off || !edges || !step_off || !node_off || !node_map) {
    set_error("euler_paths: null pointer / empty batch");
    return -1;
  }
  std::vector<std::vector<int>> all_steps(n_graphs), all_maps(n_graphs);
  std::atomic<int> next(0);
  auto work = [&]() {
    for (int g = next.fetch_add(1); g < n_graphs; g = next.fetch_add(1)) {
      const long long e0 = edge_off[g], e1 = edge_off[g + 1];
      graph_walk(node_count[g], edges + 2 * e0, static_cast<int>(e1 - e0), seed * 0x9E3779B97F4A7C15ull + g, scope,
                 all_steps[g], all_maps[g]);
    }
  };
  int nt = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  if (nt < 1) nt = 1;
  if (nt > n_graphs) nt = n_graphs;
  std::vector<std::thread> pool;
  for (int i = 1; i < nt; ++i) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  long long total = 0, nodes = 0;
  for (int g = 0; g < n_graphs; ++g) {
    step_off[g] = total;
    node_off[g] = nodes;
    total += static_cast<long long>(all_steps[g].size()) / 3;
    nodes += node_count[g];
  }
  step_off[n_graphs] = total;
  node_off[n_graphs] = nodes;
  if (steps == nullptr || total > steps_cap) return total;   // sizing call: the caller allocates 3 * total ints and repeats
  for (int g = 0; g < n_graphs; ++g) {
    if (!all_steps[g].empty()) std::memcpy(steps + 3 * step_off[g], all_steps[g].data(), all_steps[g].size() * sizeof(int));
    std::memcpy(node_map + node_off[g], all_maps[g].data(), all_maps[g].size() * sizeof(int));
  }
  return total;
}

int ggpt_stack_path_rows(const int* row_node, const int* row_edge, const int* node_map, const long long* node_attr,
                         int n_node_attr, const long long* edge_attr, int n_edge_attr, const long long* default_edge,
                         long long node_base, long long* rows, long long R, void* stream) {
  GGPT_REQUIRE(row_node && row_edge && node_map && rows && R > 0, "stack_path_rows: null pointer / empty");
  GGPT_REQUIRE(n_node_attr >= 0 && n_edge_attr >= 0 && (n_node_attr == 0 || node_attr) &&
                   (n_edge_attr == 0 || (edge_attr && default_edge)),
               "stack_path_rows: attribute tables missing");
  const long long total = R * (1 + n_node_attr + n_edge_attr);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  stack_path_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      row_node, row_edge, node_map, node_attr, n_node_attr, edge_attr, n_edge_attr, default_edge, node_base, rows, R);
  return check_launch("stack_path_rows_kernel");
}

}  // extern "C"
