"""tests/golden/aux/packing.pt: outputs of the REFERENCE's `pack_token_seq` (src/data/tokenizer.py:359-415), called as an
unbound method on a minimal stand-in for the dataset-bound tokenizer (only the attributes the method touches), plus the
block-diagonal mask built exactly as src/utils/tokenizer_utils.py:228-233,246,351-355 do from its `ls_len`.
Build container only.  Re-run:  python tests/golden/make_golden_packing.py"""
import os
import random
import sys

import numpy as np
import torch
from scipy.linalg import block_diag

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def main():
    ref_shim._Finder.roots = tuple(r for r in ref_shim._Finder.roots if r != "accelerate")
    ref_shim.load_reference()
    from torch.utils.data import Dataset
    from src.data import tokenizer as rtok
    from src.utils import tokenizer_utils as tu
    recs = []
    for case, (F_, mpe, seed) in enumerate([(5, 64, 5), (1, 48, 6), (13, 128, 7)]):
        rng = np.random.default_rng(seed)
        lens = rng.integers(3, 30, size=50)
        if F_ == 1:
            graphs = [rng.integers(22, 700, size=int(L)).tolist() for L in lens]
            eos_row = [19]
        else:
            graphs = [rng.integers(22, 700, size=(int(L), F_)).tolist() for L in lens]
            eos_row = [19] * F_

        class DS(Dataset):
            def __len__(self):
                return len(graphs)

            def __getitem__(self, i):
                return i, i

        class Fake:
            pass

        fake = Fake()
        fake.dataset, fake.sampler, fake.mpe, fake.random_ratio, fake.token_components = DS(), list(range(len(graphs))), mpe, 1.0, None
        fake.get_token_components = lambda ls, fake=fake: rtok.GSTTokenizer.get_token_components(fake, ls)
        fake.get_eos_token = lambda: 19
        fake.get_label_pad_token = lambda: -100
        fake.tokenize = lambda g, graphs=graphs: tu.TokenizationOutput(ls_tokens=[r if F_ == 1 else list(r) for r in graphs[g]],
                                                                     ls_labels=[r if F_ == 1 else list(r) for r in graphs[g]], ls_embed=[])
        first = 7
        random.seed(100 + case)
        ls_tokens, ls_labels, ls_embed, ls_len = rtok.GSTTokenizer.pack_token_seq(fake, fake.tokenize(first), first)
        random.seed(100 + case)                       # replay Python's random stream: which graphs were appended
        order, tot = [first], len(graphs[first]) + 1
        while tot < mpe:
            random.uniform(0, 1.0)
            g = random.choice(fake.sampler)
            order.append(g)
            tot += len(graphs[g]) + 1
        # prepare_inputs_for_pretrain_mlm: final <eos> row (:228-233), ls_len[-1] += 1 (:246), block_diag (:351-355);
        # collator pad(): truncate to pad_to = mpe (tokenizer.py:340-357)
        input_ids = ls_tokens + [eos_row if F_ > 1 else 19]
        ls_len = list(ls_len)
        ls_len[-1] += 1
        blocks = np.array(ls_len) - np.array([0] + ls_len[:-1])
        am = block_diag(*[np.ones([b, b], dtype=int) for b in blocks])[:mpe, :mpe]
        recs.append(dict(F=F_, mpe=mpe, graphs=graphs, order=order, sep=eos_row,
                         input_ids=torch.tensor(input_ids[:mpe]).reshape(min(len(input_ids), mpe), -1),
                         attention_mask=torch.from_numpy(np.ascontiguousarray(am)), ls_len=ls_len))
        print(f"case {case}: F={F_} mpe={mpe} graphs {order} rows {len(input_ids)}")
    torch.save(recs, os.path.join(HERE, "aux", "packing.pt"))


if __name__ == "__main__":
    main()
