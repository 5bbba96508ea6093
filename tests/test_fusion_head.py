"""TQN fusion head + DQNCOSLoss: the oracle restatement and the drop-in's host logic
against fixtures the REFERENCE produced (tests/golden/fusion.npz, oracle/make_golden.py:golden_fusion).  On CPU the
kernels are the contract model (tests/kernel_model.py); the GPU variants at the bottom run the real K4 / LSE kernels and
run on every GPU test pass (validated on a B200 in round 2)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import oracle  # noqa: E402
import xtag_clip_b200 as xt  # noqa: E402
from kernel_model import ModelKernels  # noqa: E402


def rel_err(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "fusion.npz"), allow_pickle=False)


def _case(G, n):
    pre = f"f{n}_"
    seed, d, layers, B, Pi, Pt = (int(v) for v in G[pre + "cfg"])
    return pre, seed, d, layers


def test_oracle_matches_reference_fusion(G):
    for n in range(int(G["n_cases"])):
        pre, seed, d, layers = _case(G, n)
        params = oracle.make_fusion_params(seed, d, 1, layers)
        assert sorted(params.keys()) == list(G[pre + "state_keys"])         # the reference's state_dict names
        img = torch.from_numpy(G[pre + "out_token"]).requires_grad_(True)
        txt = torch.from_numpy(G[pre + "text_tokens"]).requires_grad_(True)
        i2t = oracle.fusion_scores(img, txt, params, layers)
        t2i = oracle.fusion_scores(txt, img, params, layers)
        assert rel_err(i2t, G[pre + "i2t"]) < 1e-11 and rel_err(t2i, G[pre + "t2i"]) < 1e-11
        l1, l2 = oracle.dqn_cos_loss(i2t), oracle.dqn_cos_loss(t2i)
        assert abs(float(l1) - float(G[pre + "loss_i2t"])) < 1e-12 and abs(float(l2) - float(G[pre + "loss_t2i"])) < 1e-12
        (l1 + l2).backward()
        assert rel_err(img.grad, G[pre + "d_out_token"]) < 1e-10 and rel_err(txt.grad, G[pre + "d_text_tokens"]) < 1e-10


def test_oracle_dqn_cos_loss_wide_range(G):
    for n in range(3):
        x = torch.from_numpy(G[f"ce{n}_x"]).requires_grad_(True)
        l = oracle.dqn_cos_loss(x)
        l.backward()
        assert abs(float(l) - float(G[f"ce{n}_loss"])) <= 1e-12 * max(1.0, abs(float(l)))
        assert rel_err(x.grad, G[f"ce{n}_dx"]) < 1e-12


def _head(seed, d, layers, kernels, dtype=torch.float64, device="cpu"):
    head = xt.FusionHead(d, 1, layers, _kernels=kernels)
    sd = oracle.make_fusion_params(seed, d, 1, layers)
    head.load_state_dict(sd, strict=True)                                   # reference key names, incl. the inert ones
    return head.to(device=device, dtype=dtype).eval()


def test_fusion_head_host_logic_matches_reference(G):
    """FusionHead / fusion_scores / DQNCOSLoss with the contract-model kernels: values, token gradients and two
    weight-gradient slices equal what the reference's TQN_Model + DQNCOSLoss produced."""
    for n in range(int(G["n_cases"])):
        pre, seed, d, layers = _case(G, n)
        K = ModelKernels()
        head = _head(seed, d, layers, K)
        assert sorted(head.state_dict().keys()) == list(G[pre + "state_keys"])
        img = torch.from_numpy(G[pre + "out_token"]).requires_grad_(True)
        txt = torch.from_numpy(G[pre + "text_tokens"]).requires_grad_(True)
        i2t = xt.fusion_scores(head, img, txt)
        t2i = xt.fusion_scores(head, txt, img)
        assert rel_err(i2t, G[pre + "i2t"]) < 1e-9 and rel_err(t2i, G[pre + "t2i"]) < 1e-9
        ce = xt.DQNCOSLoss(_kernels=K)
        l1, l2 = ce(i2t), ce(t2i)
        assert abs(float(l1) - float(G[pre + "loss_i2t"])) < 1e-6 and abs(float(l2) - float(G[pre + "loss_t2i"])) < 1e-6
        (l1 + l2).backward()
        assert rel_err(img.grad, G[pre + "d_out_token"]) < 1e-5 and rel_err(txt.grad, G[pre + "d_text_tokens"]) < 1e-5
        gi = head.decoder.layers[0].multihead_attn.in_proj_weight.grad[:6, :8]
        assert rel_err(gi, G[pre + "d_inproj0"]) < 1e-5
        assert rel_err(head.mlp_head[9].weight.grad, G[pre + "d_mlp9"]) < 1e-5


def test_fusion_head_query_chunks_and_seq_first_queries():
    """More queries than one K4 launch holds (chunks of 64 share K/V), and the `inside_repeat=False` calling form."""
    K = ModelKernels()
    d, layers, B, P, Q = 32, 2, 3, 11, 150
    head = _head(5, d, layers, K)
    params = oracle.make_fusion_params(5, d, 1, layers)
    g = torch.Generator().manual_seed(3)
    mem = torch.randn(B, P, d, generator=g, dtype=torch.float64)
    qf = torch.randn(Q, d, generator=g, dtype=torch.float64)
    out = head(mem, qf)
    assert out.shape == (B, Q, 1)
    assert rel_err(out, oracle.fusion_forward(mem, qf, params, layers)) < 1e-9
    assert sum(1 for c in K.calls if c[0] == "xattn_fwd") == layers * 3     # ceil(150 / 64) launches per layer
    out2 = head(mem, qf.unsqueeze(1).repeat(1, B, 1), inside_repeat=False)
    assert rel_err(out2, out) < 1e-12
    with pytest.raises(NotImplementedError):
        head(mem, qf, return_atten=True)
    with pytest.raises(ValueError):
        head(mem[..., :16], qf)
    with pytest.raises(ValueError):
        xt.DQNCOSLoss(_kernels=K)(torch.zeros(3, 4))


def test_dqn_cos_loss_drop_in_matches_reference(G):
    K = ModelKernels()
    for n in range(3):
        x = torch.from_numpy(G[f"ce{n}_x"]).float().requires_grad_(True)
        l = xt.DQNCOSLoss(_kernels=K)(x)
        l.backward()
        # fp32 LSEs of logits up to |x| ~ 400: the exponent rounding (eps * |x|) is the error of the probabilities
        assert abs(float(l) - float(G[f"ce{n}_loss"])) <= 2e-6 * max(1.0, abs(float(G[f"ce{n}_loss"])))
        assert rel_err(x.grad, G[f"ce{n}_dx"]) < 5e-5


def test_from_reference_live():
    """Against the reference tree itself when it is present (build container only)."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ref_shim.load_ref_open_clip()
    from open_clip.CAR_heads.TQN_model import TQN_Model
    torch.manual_seed(0)
    ref = TQN_Model().double().eval()
    head = xt.FusionHead.from_reference(ref, _kernels=ModelKernels()).eval()
    g = torch.Generator().manual_seed(9)
    mem = torch.randn(2, 8, 512, generator=g, dtype=torch.float64)
    qf = torch.randn(2, 512, generator=g, dtype=torch.float64)
    assert rel_err(head(mem, qf), ref(mem, qf)) < 1e-9


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 3e-2)])
def test_fusion_head_gpu(G, dtype, tol):
    pre, seed, d, layers = _case(G, 2)                                      # d_model 512: head dim 128 (tensor-core K4)
    head = _head(seed, d, layers, None, dtype=dtype, device="cuda")
    img = torch.from_numpy(G[pre + "out_token"]).to("cuda", dtype).requires_grad_(True)
    txt = torch.from_numpy(G[pre + "text_tokens"]).to("cuda", dtype).requires_grad_(True)
    i2t = xt.fusion_scores(head, img, txt)
    assert rel_err(i2t.float(), G[pre + "i2t"]) < tol
    loss = xt.DQNCOSLoss()(i2t.float())
    assert abs(float(loss) - float(G[pre + "loss_i2t"])) < tol
    loss.backward()
    assert torch.isfinite(img.grad.float()).all() and torch.isfinite(txt.grad.float()).all()


@pytest.mark.gpu
def test_dqn_cos_loss_gpu(G):
    for n in range(3):
        x = torch.from_numpy(G[f"ce{n}_x"]).float().cuda().requires_grad_(True)
        l = xt.DQNCOSLoss()(x)
        l.backward()
        assert abs(float(l) - float(G[f"ce{n}_loss"])) <= 1e-5 * max(1.0, abs(float(G[f"ce{n}_loss"])))
        assert rel_err(x.grad, G[f"ce{n}_dx"]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("n,dtype", [(7, torch.float32), (130, torch.float32), (1024, torch.bfloat16), (333, torch.bfloat16)])
def test_symm_ce_kernels_gpu(n, dtype):
    """csrc/symm_ce.cu against the oracle's DQNCOSLoss on random matrices with a wide range (|x| up to ~60), ragged
    sizes and a strided (non-contiguous rows) input."""
    g = torch.Generator().manual_seed(n)
    big = torch.randn(n, n + 8, generator=g) * 20
    x = big[:, :n].to(dtype).cuda().requires_grad_(True)               # row stride n + 8 after .cuda()? keep it strided:
    xs = big.to(dtype).cuda()[:, :n].detach().requires_grad_(True)
    for inp in (x, xs):
        loss = xt.DQNCOSLoss()(inp)
        (loss * 1.7).backward()
        x64 = inp.detach().double().cpu().requires_grad_(True)
        ref = oracle.dqn_cos_loss(x64)
        (ref * 1.7).backward()
        assert abs(float(loss) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
        assert rel_err(inp.grad.float(), x64.grad) < (1e-5 if dtype == torch.float32 else 6e-3)


@pytest.mark.gpu
def test_fusion_head_long_query_set_gpu():
    """B = 96 samples -> every sample attends with Lq = 96 queries (> 64): the bf16 path runs K4 in ONE launch per
    layer (forward chunks on the grid, two-launch backward), the exact fp32 path in 64-query chunks; both against the
    oracle (fp64) on the same inputs.  Scores: 2e-4 (fp32) / 3e-2 (bf16) max-norm.  Token gradients: the ReLU MLP makes
    them discontinuous in the inputs (a CPU fp32 run of the same head already differs from fp64 by 3e-3 max-norm), so
    they are held to a direction / magnitude bar (cosine >= 0.999 / 0.98, norm within 1 % / 10 %); the kernels
    themselves are pinned tightly in test_gpu_kernels.py::test_xattn_long_query_sets_one_launch."""
    torch.backends.cuda.matmul.allow_tf32 = False      # projections / MLP are torch library calls
    seed, d, layers, B, Pi, Pt = 70, 512, 2, 96, 9, 7
    params = oracle.make_fusion_params(seed, d, 1, layers)
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, Pi, d, generator=g) * 0.5
    txt = torch.randn(B, Pt, d, generator=g) * 0.5
    wgt = torch.randn(B, B, generator=g)
    i64 = img.double().requires_grad_(True)
    t64 = txt.double().requires_grad_(True)
    ref = oracle.fusion_scores(i64, t64, params, layers)
    (ref * wgt.double()).sum().backward()

    def direction(a, b):
        a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
        return float(a @ b / (a.norm() * b.norm())), float(a.norm() / b.norm())

    for dtype, tol_s, cos_min, norm_tol in ((torch.float32, 2e-4, 0.999, 0.01), (torch.bfloat16, 3e-2, 0.98, 0.10)):
        head = _head(seed, d, layers, None, dtype=dtype, device="cuda")
        ic = img.to("cuda", dtype).requires_grad_(True)
        tc = txt.to("cuda", dtype).requires_grad_(True)
        out = xt.fusion_scores(head, ic, tc)
        (out.float() * wgt.cuda()).sum().backward()
        assert rel_err(out.float(), ref) < tol_s, (dtype, rel_err(out.float(), ref))
        for got, want in ((ic.grad, i64.grad), (tc.grad, t64.grad)):
            c, r = direction(got.float(), want)
            assert c >= cos_min and abs(r - 1.0) <= norm_tol, (dtype, c, r)
