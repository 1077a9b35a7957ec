"""CPU oracle for sequence packing — TEST INFRASTRUCTURE, NOT PRODUCT CODE (rules as in graphgpt_oracle.py).

Restates, with plain Python lists and numpy as the reference does, what its tokenizer produces for one packed sample:
  pack_token_seq (src/data/tokenizer.py:359-415): ls_tokens = g0 + [sep] + g1 + [sep] + ... ; ls_len = running lengths
  prepare_inputs_for_pretrain_mlm (src/utils/tokenizer_utils.py:228-233,246): one more <eos> row after the last graph
  block-wise attention (tokenizer_utils.py:351-355): scipy-style block_diag of ones, one block per ls_len increment
  collator pad() (tokenizer.py:340-357): truncate to pad_to = mpe, right-pad.
PARITY PIN: pinned.  tests/golden/make_golden_packing.py calls the reference's own `pack_token_seq` (as an unbound
method on a stand-in exposing only the attributes it touches; F = 1, 5, 13) and builds the mask from its `ls_len` with
the reference's block_diag lines; tests/test_oracle_golden.py::test_packing_oracle_matches_reference_pack_token_seq
compares token rows, mask and segment boundaries exactly.  tests/test_host_cpu.py adds an independent scipy check."""
import numpy as np


def pack_one(graph_rows, sep_row, mpe, pad_id=0):
    """graph_rows: list of int arrays [len_g, F].  Returns (input_ids [mpe,F], attention_mask [mpe,mpe], seg_ids [mpe])."""
    F = len(sep_row)
    ls_tokens, ls_len = [], []
    for g in graph_rows:
        ls_tokens.extend([list(r) for r in g])
        ls_tokens.append(list(sep_row))                 # separator between graphs / the final <eos> row
        ls_len.append(len(ls_tokens))
    lens = np.array(ls_len) - np.array([0] + ls_len[:-1])
    total = int(lens.sum())
    am = np.zeros((total, total), dtype=np.int64)
    seg = np.zeros((total,), dtype=np.int64)
    o = 0
    for k, L in enumerate(lens):                        # block_diag(*[ones(L, L)])
        am[o:o + L, o:o + L] = 1
        seg[o:o + L] = k + 1
        o += L
    ids = np.full((mpe, F), pad_id, dtype=np.int64)
    n = min(total, mpe)
    ids[:n] = np.asarray(ls_tokens[:n], dtype=np.int64).reshape(n, F)
    am_out = np.zeros((mpe, mpe), dtype=np.int64)
    am_out[:n, :n] = am[:n, :n]
    seg_out = np.zeros((mpe,), dtype=np.int64)
    seg_out[:n] = seg[:n]
    return ids, am_out, seg_out
