"""Synthetic GraphGPT batches shaped like the reference's DataCollatorForGST output.

Host-side numpy/networkx only; it stands in for the reference's CPU data pipeline (src/data/collator.py:70-111,
src/data/tokenizer.py:994-1139, src/utils/tokenizer_utils.py:224-363) which is out of scope for the hot path.

A sample is an (eulerised) Eulerian path over a small molecule-like graph, serialised as rows of F stacked tokens:
  column 0      node-id token, cyclic re-index from a random start (nx_utils.py:234-260)
  columns 1..   node-attribute tokens, then edge-attribute tokens (edge leading INTO the row's node;
                row 0 carries the default edge-attribute tokens)
  last row      all <eos> (tokenizer_utils.py:230)
SMTP masking: per sample t~U(0.01,0.99), every non-pad entry masked w.p. 1-t -> input <mask>=1, label = original id,
else label -100 (tokenizer_utils.py:112-172,256-269).
"""
from dataclasses import dataclass

import numpy as np

MASK_ID, EOS_ID, BOS_ID, PAD_ID = 1, 19, 20, 0
NODE_ID_BASE = 22


@dataclass
class VocabLayout:
    vocab_size: int = 756
    scope: int = 512          # node-id tokens [22, 22+scope)
    n_node_attr: int = 9
    n_edge_attr: int = 3

    @property
    def stacked_feat(self):
        return 1 + self.n_node_attr + self.n_edge_attr

    def attr_ranges(self):
        """Disjoint id sub-range per attribute column inside [attr_base, vocab_size)."""
        attr_base = NODE_ID_BASE + self.scope + 23  # 23 reserved / number tokens (vocab_builder.py:113-131)
        ncol = self.n_node_attr + self.n_edge_attr
        if ncol == 0:
            return []
        width = max(1, (self.vocab_size - attr_base) // ncol)
        out = []
        for c in range(ncol):
            lo = attr_base + c * width
            out.append((min(lo, self.vocab_size - 1), min(lo + width, self.vocab_size)))
        return out


PCQM_VOCAB = VocabLayout()
TOY_VOCAB = VocabLayout(vocab_size=300, scope=128, n_node_attr=0, n_edge_attr=0)
PPA_VOCAB = VocabLayout(vocab_size=41244, scope=512, n_node_attr=2, n_edge_attr=1)  # F = 1+2+1 = 4


def _euler_walk(n_nodes, n_rings, rng):
    """Node sequence of an Eulerian walk over a random bounded-degree tree + ring closures, eulerised by doubling
    the tree edges on the way back (a DFS walk that drops the final return, as shorten_path does,
    nx_utils.py:331-348).  networkx-free so it runs anywhere; calibrated against the networkx simulation in
    SURVEY §8d (mean ~23.5 rows)."""
    parent = [-1] * n_nodes
    deg = [0] * n_nodes
    children = [[] for _ in range(n_nodes)]
    for v in range(1, n_nodes):
        while True:
            u = int(rng.integers(max(0, v - 6), v))
            if deg[u] < 3:
                break
            if all(deg[w] >= 3 for w in range(max(0, v - 6), v)):
                u = int(np.argmin(deg[:v]))
                break
        parent[v] = u
        deg[u] += 1
        deg[v] += 1
        children[u].append(v)
    extra = {}
    for _ in range(n_rings):
        a, b = (int(x) for x in rng.integers(0, n_nodes, 2))
        if a != b and parent[a] != b and parent[b] != a:
            extra.setdefault(a, []).append(b)
            extra.setdefault(b, []).append(a)
    walk = []
    used = set()

    def dfs(u):
        walk.append(u)
        for w in extra.get(u, []):
            key = (min(u, w), max(u, w))
            if key not in used:  # ring-closure edge: go there and come straight back
                used.add(key)
                walk.append(w)
                walk.append(u)
        order = list(children[u])
        rng.shuffle(order)
        for c in order:
            dfs(c)
            walk.append(u)

    dfs(0)
    # shorten_path: stop right after the last step that covers a not-yet-visited edge
    seen, last_new = set(), 0
    for i in range(1, len(walk)):
        key = (min(walk[i - 1], walk[i]), max(walk[i - 1], walk[i]))
        if key not in seen:
            seen.add(key)
            last_new = i
    walk = walk[: last_new + 1]
    return walk


def sample_graph_rows(rng, vocab: VocabLayout = PCQM_VOCAB):
    """One tokenised sample: int64 [rows, F] (before SMTP masking), rows = |walk| + 1 (<eos> row)."""
    n_nodes = int(np.clip(np.rint(rng.normal(12.4, 2.4)), 2, 20))
    n_rings = int(rng.integers(0, 3))
    walk = _euler_walk(n_nodes, n_rings, rng)
    F = vocab.stacked_feat
    rows = np.zeros((len(walk) + 1, F), dtype=np.int64)
    start = int(rng.integers(0, vocab.scope))
    first_seen = {}
    for v in walk:
        if v not in first_seen:
            first_seen[v] = len(first_seen)
    rows[:-1, 0] = [NODE_ID_BASE + (start + first_seen[v]) % vocab.scope for v in walk]
    ranges = vocab.attr_ranges()
    node_attr = {}
    for c in range(vocab.n_node_attr):
        lo, hi = ranges[c]
        for v in first_seen:
            node_attr[(v, c)] = int(rng.integers(lo, hi))
        rows[:-1, 1 + c] = [node_attr[(v, c)] for v in walk]
    for c in range(vocab.n_edge_attr):
        lo, hi = ranges[vocab.n_node_attr + c]
        col = 1 + vocab.n_node_attr + c
        rows[0, col] = lo  # default edge-attr token on the first row (tokenizer.py:1024-1027)
        if len(walk) > 1:
            rows[1:-1, col] = rng.integers(lo, hi, size=len(walk) - 1)
    rows[-1, :] = EOS_ID
    return rows


def smtp_mask(rows, rng):
    """Entry-level scheduled masking (tokenizer_utils.py:112-148,256-269).  Returns (input_ids, labels, t)."""
    t = float(rng.uniform(0.01, 0.99))
    m = rng.random(rows.shape) < (1.0 - t)
    labels = np.where(m, rows, -100)
    inputs = np.where(m & (rows != PAD_ID), MASK_ID, rows)
    return inputs, labels, t


def ntp_labels(rows):
    """Causal next-row prediction labels: row i predicts row i+1, last row predicts <eos>."""
    labels = np.full_like(rows, EOS_ID)
    labels[:-1] = rows[1:]
    return labels


def make_batch(n_seq, seq_len, *, layout="packed", task="smtp", vocab: VocabLayout = PCQM_VOCAB, seed=1234,
               return_segments=False):
    """Build one batch dict of numpy arrays with the collator's keys.

    layout = "packed"   : graphs concatenated up to seq_len (last one truncated), block-diagonal
                          attention_mask [N,S,S] (tokenizer.py:359-415, tokenizer_utils.py:351-355)
             "unpacked" : one graph per row, right-padded to the batch max rounded up to 8 (tokenizer.py:627-636),
                          attention_mask [N,S]
             "dense"    : like packed but ONE attention span per row (attention_mask [N,S] all ones) — the
                          roofline workload where every query sees all seq_len keys
    """
    rng = np.random.default_rng(seed)
    F = vocab.stacked_feat
    samples = []
    seg_lens = []
    if layout == "unpacked":
        raw = [sample_graph_rows(rng, vocab) for _ in range(n_seq)]
        S = min(seq_len, (max(r.shape[0] for r in raw) + 7) // 8 * 8)
        ids = np.zeros((n_seq, S, F), np.int64)
        labels = np.full((n_seq, S, F), -100, np.int64)
        am = np.zeros((n_seq, S), np.int64)
        for i, r in enumerate(raw):
            r = r[:S]
            if task == "smtp":
                inp, lab, _ = smtp_mask(r, rng)
            else:
                inp, lab = r, ntp_labels(r)
            ids[i, : len(r)] = inp
            labels[i, : len(r)] = lab
            am[i, : len(r)] = 1
            seg_lens.append([len(r)])
        pos = np.broadcast_to(np.arange(S, dtype=np.int64), (n_seq, S)).copy()
        batch = {"input_ids": ids, "labels": labels, "attention_mask": am, "position_ids": pos}
    else:
        S = seq_len
        ids = np.zeros((n_seq, S, F), np.int64)
        labels = np.full((n_seq, S, F), -100, np.int64)
        for i in range(n_seq):
            fill = 0
            lens = []
            while fill < S:
                r = sample_graph_rows(rng, vocab)[: S - fill]
                if task == "smtp":
                    inp, lab, _ = smtp_mask(r, rng)
                else:
                    inp, lab = r, ntp_labels(r)
                ids[i, fill: fill + len(r)] = inp
                labels[i, fill: fill + len(r)] = lab
                fill += len(r)
                lens.append(len(r))
            seg_lens.append(lens)
        pos = np.broadcast_to(np.arange(S, dtype=np.int64), (n_seq, S)).copy()
        if layout == "packed":
            am = np.zeros((n_seq, S, S), np.int64)
            for i, lens in enumerate(seg_lens):
                o = 0
                for L in lens:
                    am[i, o: o + L, o: o + L] = 1
                    o += L
        elif layout == "dense":
            am = np.ones((n_seq, S), np.int64)
        else:
            raise ValueError(layout)
        batch = {"input_ids": ids, "labels": labels, "attention_mask": am, "position_ids": pos}
    if F == 1:
        batch["input_ids"] = batch["input_ids"][:, :, 0]
        batch["labels"] = batch["labels"][:, :, 0]
    if return_segments:
        batch["segment_lens"] = seg_lens
    return batch


def make_ft_batch(n_seq, seq_len, *, vocab: VocabLayout = PPA_VOCAB, seed=1234, min_frac=0.25, embed_dim=0, num_labels=2):
    """Edge-level fine-tuning batch shaped like ft_batch_training's input (training_utils.py:98-205): right-padded
    [N,S] sequences (sub-graph samples of varying length, the longest one filling seq_len), `position_ids`,
    `edge_labels` ~ Bernoulli(0.5) (tokenizer_utils.py:571-633 appends the two target-node rows; here they are just the
    last two valid rows), `labels` all -100 (no auxiliary LM loss), optional raw node embeddings `embed` [N,S,E]."""
    rng = np.random.default_rng(seed + 7)
    b = make_batch(n_seq, seq_len, layout="dense", task="ntp", vocab=vocab, seed=seed)
    lens = rng.integers(max(2, int(seq_len * min_frac)), seq_len + 1, size=n_seq)
    lens[0] = seq_len
    am = (np.arange(seq_len)[None, :] < lens[:, None]).astype(np.int64)
    ids = b["input_ids"].copy()
    ids[am == 0] = PAD_ID
    out = {"input_ids": ids, "attention_mask": am, "position_ids": b["position_ids"],
           "labels": np.full_like(ids, -100), "edge_labels": rng.integers(0, num_labels, size=n_seq).astype(np.int64),
           "segment_lens": [[int(l)] for l in lens]}
    if embed_dim > 0:
        emb = rng.standard_normal((n_seq, seq_len, embed_dim)).astype(np.float32)
        emb[am == 0] = 0.0
        out["embed"] = emb
    return out


def rows_per_sample_stats(n=3000, seed=0, vocab: VocabLayout = PCQM_VOCAB):
    rng = np.random.default_rng(seed)
    lens = np.array([sample_graph_rows(rng, vocab).shape[0] for _ in range(n)])
    return {"mean": float(lens.mean()), "std": float(lens.std()), "p5": float(np.percentile(lens, 5)),
            "p50": float(np.percentile(lens, 50)), "p95": float(np.percentile(lens, 95)), "max": int(lens.max())}
