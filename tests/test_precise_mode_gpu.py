"""The validation-only PRECISE mode (GGPT_PRECISE=1, graph-gpt_b200/precise.py + csrc/precise.cu) against the fp32
goldens of the unmodified reference and against the fp32 oracle at full depth.

The north star's tolerance is 1e-3 relative on logits AND loss.  The fast path stores bf16 and lands at 3..6e-3 on the
logits (tests/test_parity_depth_gpu.py explains that number); this mode sends every GEMM through the SAME tcgen05 kernel
on split-bf16 operands with fp32 activations in between, and must meet 1e-3 outright — measured ~1e-5, i.e. the GEMM
kernel, the mask plan, the label compaction and the loss are arithmetically exact and the fast path's distance is bf16
storage rounding.  Numbers are appended to gpurun_out/parity_depth_report.txt."""
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))
REPORT = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out", "parity_depth_report.txt")
TOL = 1e-3          # BASELINE.json north_star: "logits and loss matching the reference within 1e-3 relative"


def _relf(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _log(msg):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(msg + "\n")
    print(msg)


@pytest.fixture
def precise_env(monkeypatch):
    monkeypatch.setenv("GGPT_PRECISE", "1")


def test_split_operand_gemm_is_fp32_accurate(precise_env):
    """precise.linear (tcgen05 GEMM over K' = 3K split-bf16 operands) vs an fp64 matmul: ~2^-17, ragged M / N / K."""
    from graphgpt_b200 import precise
    g = torch.Generator(device="cuda").manual_seed(0)
    for M, N, K in [(300, 756, 768), (129, 72, 64), (1024, 2304, 768), (7, 9984, 768)]:
        x = torch.randn((M, K), device="cuda", generator=g)
        w = torch.randn((N, K), device="cuda", generator=g) * 0.05
        y = precise.linear(x, w)
        ref = x.double() @ w.double().t()
        e = _relf(y, ref)
        assert tuple(y.shape) == (M, N) and e <= 2e-5, (M, N, K, e)


def test_precise_attention_matches_fp64(precise_env):
    from graphgpt_b200 import ops
    from graphgpt_b200.lib import lib
    N, S, H = 2, 200, 3
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv = torch.randn((N * S, 3 * d), device="cuda", generator=g)
    am = torch.ones((N, S), dtype=torch.int64, device="cuda")
    am[1, 150:] = 0
    for causal in (False, True):
        mask = ops.attn_mask_build(am, N, S, causal, qkv.device)
        out = torch.empty((N * S, d), device="cuda")
        lib.ggpt_vp_attn_f32(qkv.data_ptr(), 3 * d, 0, d, 2 * d, mask.bits.data_ptr(), out.data_ptr(), d, N, S, H,
                             torch.cuda.current_stream().cuda_stream)
        q, k, v = (qkv[:, i * d:(i + 1) * d].double().view(N, S, H, 64).transpose(1, 2) for i in range(3))
        keep = am[:, None, None, :].bool().expand(N, 1, S, S).clone()
        if causal:
            keep &= torch.ones((S, S), dtype=torch.bool, device="cuda").tril()
        sc = (q @ k.transpose(2, 3)) * 0.125
        sc = sc.masked_fill(~keep, float("-inf"))
        ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(N * S, d)
        valid = am.reshape(-1).bool()
        assert _relf(out[valid], ref[valid]) <= 1e-5


def _build(rec):
    from graphgpt_b200 import GraphGPTConfig, GraphGPTDoubleHeadsModel, GraphGPTPretrainBase, GraphGPTTaskModel
    cfg = GraphGPTConfig(**rec["config"])
    cls = {"pretrain": GraphGPTPretrainBase, "finetune": GraphGPTTaskModel, "double": GraphGPTDoubleHeadsModel}[rec["kind"]]
    model = cls(cfg)
    missing, unexpected = model.load_state_dict(rec["state_dict"], strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return model.cuda().eval()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_precise_mode_meets_1e3_on_reference_goldens(path, precise_env):
    rec = torch.load(path)
    model = _build(rec)
    inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in rec["inputs"].items()}
    name = os.path.basename(path)[:-3]
    with torch.no_grad():
        try:
            out = model(**inp)
        except NotImplementedError as e:          # token_ce_intra / [N,S,S,E] raw embeddings: fast path only
            pytest.skip(str(e))
    if rec["kind"] == "pretrain":
        e = _relf(out.head1_logits[:: rec["logits_stride"]], rec["logits"])
        msg = f"precise {name}: head1_logits relF {e:.3e}"
        assert tuple(out.head1_logits.shape) == tuple(rec["logits_shape"]) and e <= TOL, e
        if "loss" in rec:
            el = abs(out.head1_loss.item() - rec["loss"].item()) / abs(rec["loss"].item())
            msg += f" loss rel {el:.3e}"
            assert el <= TOL, el
    elif rec["kind"] == "double":
        e = _relf(out.task_logits, rec["task_logits"])
        et = abs(out.task_loss.item() - rec["task_loss"].item()) / abs(rec["task_loss"].item())
        ep = abs(out.pretrain_loss.item() - rec["pretrain_loss"].item()) / abs(rec["pretrain_loss"].item())
        msg = f"precise {name}: task_logits relF {e:.3e} task_loss rel {et:.3e} pretrain_loss rel {ep:.3e}"
        assert e <= TOL and et <= TOL and ep <= TOL, (e, et, ep)
    else:
        am = rec["inputs"]["attention_mask"].bool()
        if out.task_logits.dim() == 3:
            e = _relf(out.task_logits.float().cpu()[am], rec["task_logits"][am])
        else:
            e = _relf(out.task_logits, rec["task_logits"])
        eh = _relf(out.hidden_states.float().cpu()[am], rec["hidden"][am])
        eth = _relf(out.task_hidden_states, rec["task_hidden"])
        msg = f"precise {name}: task_logits relF {e:.3e} hidden relF {eh:.3e} task_hidden relF {eth:.3e}"
        assert e <= TOL and eh <= TOL and eth <= TOL, (e, eh, eth)
        if "loss" in rec:
            el = abs(out.task_loss.item() - rec["loss"].item()) / abs(rec["loss"].item())
            msg += f" loss rel {el:.3e}"
            assert el <= TOL, el
    _log(msg)


def test_precise_mode_full_depth_c2_vs_fp32_oracle(precise_env):
    """12L/768d/F13/V756, 2 x 1024 packed: the precise mode against the fp32 CPU oracle at the headline depth."""
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase, synth
    from oracle import graphgpt_oracle as oracle
    cfgd = dict(vocab_size=756, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                num_key_value_heads=12, head_dim=64, hidden_act="gelu", max_position_embeddings=1024, rms_norm_eps=1e-6,
                rope_theta=10000.0, pad_token_id=0, bos_token_id=20, eos_token_id=19, causal_attention=False,
                stacked_feat=13, stack_method="short", stacked_feat_agg_method="sum", next_n_token=13, use_cache=False)
    b = synth.make_batch(2, 1024, layout="packed", seed=9)
    ids, am, labels = (torch.from_numpy(b[k]) for k in ("input_ids", "attention_mask", "labels"))
    sd = oracle.init_state_dict(cfgd, seed=5)
    with torch.no_grad():
        ref = oracle.pretrain_forward(sd, cfgd, ids, am, labels)
        model = GraphGPTPretrainBase(GraphGPTConfig(**cfgd))
        model.load_state_dict(sd, strict=True)
        model = model.cuda().eval()
        out = model(input_ids=ids.cuda(), attention_mask=am.cuda(), labels=labels.cuda())
    e = _relf(out.head1_logits, ref["logits"])
    el = abs(out.head1_loss.item() - ref["loss"].item()) / ref["loss"].item()
    _log(f"precise c2_12L768d_packed_2x1024: head1_logits relF {e:.3e} loss rel {el:.3e}")
    assert e <= TOL and el <= TOL, (e, el)


def test_precise_mode_refuses_training(precise_env):
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase
    cfg = GraphGPTConfig(vocab_size=300, hidden_size=64, intermediate_size=256, num_hidden_layers=1, num_attention_heads=1,
                         num_key_value_heads=1, hidden_act="gelu", stacked_feat=1, next_n_token=1, causal_attention=False)
    model = GraphGPTPretrainBase(cfg).cuda().train()
    ids = torch.randint(2, 300, (1, 16)).cuda()
    with pytest.raises(RuntimeError):
        model(input_ids=ids, attention_mask=torch.ones_like(ids), labels=ids)
