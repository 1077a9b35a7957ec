"""Device-side sequence packing: the per-batch list surgery of the reference's tokenizer
(`pack_token_seq`, src/data/tokenizer.py:359-415; final <eos> and block-diagonal mask, src/utils/tokenizer_utils.py:
228-233,351-355; truncation to `mpe` in the collator's pad(), tokenizer.py:340-357) as one kernel over a pool of
tokenised graphs that already lives in HBM.

WHICH graphs go into which packed sequence stays a host decision (the reference samples them with Python's `random`):
`plan_greedy` reproduces the reference's rule — keep appending graphs while the running length (rows + one separator
each) is below `mpe` — for a given graph order.  The packed batch carries the block-diagonal attention mask in its
[N,S] segment-id form (1-based segment per row, 0 = pad), which `GraphGPTPretrainBase.forward(attention_mask=...)`
accepts directly: 14 MB per 64 x 1024 batch instead of the 537 MB int64 [N,S,S] tensor.
"""
import numpy as np
import torch

from .lib import lib


def plan_greedy(lengths, mpe, order=None):
    """lengths[g] = token rows of graph g.  Returns (seq_graphs int32 [M], cu_seq int32 [N+1]): consecutive graphs of
    `order` are appended to a sequence while its length (rows + 1 separator per graph) is < mpe (tokenizer.py:366-371:
    `token_len = len(ls_tokens) + 1; while token_len < self.mpe`), so the last graph of a sequence may be truncated."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.arange(len(lengths)) if order is None else np.asarray(order)
    seq_graphs, cu_seq = [], [0]
    i = 0
    while i < len(order):
        tot = 0
        while i < len(order):
            seq_graphs.append(int(order[i]))
            tot += int(lengths[order[i]]) + 1
            i += 1
            if tot >= mpe:
                break
        cu_seq.append(len(seq_graphs))
    return np.asarray(seq_graphs, np.int32), np.asarray(cu_seq, np.int32)


def pack_sequences(rows, cu_rows, seq_graphs, cu_seq, seq_len, sep_row, pad_id=0):
    """rows int64 [R,F] (CUDA), cu_rows int32 [G+1], seq_graphs int32 [M], cu_seq int32 [N+1], sep_row int64 [F].
    Returns dict(input_ids [N,S,F], attention_mask [N,S] segment ids, position_ids [N,S], n_valid int32 [N])."""
    if not rows.is_cuda or rows.dtype != torch.int64 or rows.dim() != 2 or not rows.is_contiguous():
        raise RuntimeError("pack_sequences: rows must be a contiguous CUDA int64 [R,F] tensor (no CPU fallback)")
    dev = rows.device
    F_ = rows.shape[1]

    def i32(x):
        t = torch.as_tensor(x, dtype=torch.int32)
        return t.to(dev).contiguous()

    cu_rows, seq_graphs, cu_seq = i32(cu_rows), i32(seq_graphs), i32(cu_seq)
    sep = torch.as_tensor(sep_row, dtype=torch.int64).to(dev).contiguous()
    if sep.numel() != F_:
        raise RuntimeError(f"pack_sequences: sep_row has {sep.numel()} entries, rows have {F_} columns")
    N = cu_seq.numel() - 1
    ids = torch.empty((N, seq_len, F_), device=dev, dtype=torch.int64)
    seg = torch.empty((N, seq_len), device=dev, dtype=torch.int64)
    pos = torch.empty((N, seq_len), device=dev, dtype=torch.int64)
    n_valid = torch.empty((N,), device=dev, dtype=torch.int32)
    lib.ggpt_pack_sequences(rows.data_ptr(), F_, cu_rows.data_ptr(), seq_graphs.data_ptr(), cu_seq.data_ptr(), N, seq_len,
                            sep.data_ptr(), int(pad_id), ids.data_ptr(), seg.data_ptr(), pos.data_ptr(), n_valid.data_ptr(),
                            torch.cuda.current_stream().cuda_stream)
    return {"input_ids": ids, "attention_mask": seg, "position_ids": pos, "n_valid": n_valid}
