"""CPU tests of the host-side mirror of the reference API (xtag_clip_b200/{loss,tag_head,asymmetric_loss}.py).
The CUDA kernels are replaced by the contract model in tests/kernel_model.py through the `_kernels=` test hook,
so what is checked here is the *host logic*: signatures, label offsets, gradient weights, dtype handling,
state_dict compatibility and error behaviour -- against the golden fixtures produced by the reference."""
import inspect
import os
import types

import numpy as np
import pytest
import torch

import oracle
import xtag_clip_b200 as xt
from oracle.tag_oracle import make_tag_params
from kernel_model import ModelKernels


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def rel_err(a, b):
    """max-norm relative error: max|a-b| / max|b| (the measure the parity bars are stated in)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_cliploss_signature_matches_reference():
    sig = inspect.signature(xt.ClipLoss.__init__)
    names = [n for n, p in sig.parameters.items() if n != "self" and p.kind != inspect.Parameter.KEYWORD_ONLY]
    # extras (group, comm_dtype, exchange, cuda_graph, compute_dtype, ...) are keyword-only: positional calls of the
    # reference keep their meaning
    assert names == ["local_loss", "gather_with_grad", "cache_labels", "rank", "world_size", "use_horovod"]
    defaults = [sig.parameters[n].default for n in names]
    assert defaults == [False, False, False, 0, 1, False]
    fsig = inspect.signature(xt.ClipLoss.forward)
    assert list(fsig.parameters)[1:] == ["image_features", "text_features", "logit_scale", "output_dict"]
    assert fsig.parameters["output_dict"].default is False
    gsig = inspect.signature(xt.gather_features)
    assert list(gsig.parameters) == ["image_features", "text_features", "local_loss", "gather_with_grad", "rank",
                                     "world_size", "use_horovod"]


def test_cliploss_w1_golden(golden_dir):
    g = _load(golden_dir, "clip_w1.npz")
    for n in range(int(g["n_cases"])):
        pre = f"c{n}_f64_"
        I = torch.from_numpy(g[pre + "I"]).float().requires_grad_(True)
        T = torch.from_numpy(g[pre + "T"]).float().requires_grad_(True)
        s = torch.tensor(float(g[pre + "scale"]), requires_grad=True)
        loss_mod = xt.ClipLoss(_kernels=ModelKernels())
        out = loss_mod(I, T, s, output_dict=True)
        assert list(out.keys()) == ["contrastive_loss"]
        loss = out["contrastive_loss"]
        assert loss.dim() == 0 and loss.dtype == torch.float32
        loss.backward()
        # fp32 inputs/outputs against the reference's fp64 run: the BASELINE "fp32 mode" bar (1e-5)
        assert rel_err(loss.item(), g[pre + "loss"]) < 1e-5
        assert rel_err(I.grad.numpy(), g[pre + "dI"]) < 1e-5
        assert rel_err(T.grad.numpy(), g[pre + "dT"]) < 1e-5
        assert rel_err(s.grad.numpy(), g[pre + "dscale"]) < 1e-5


def test_cliploss_misc_behaviour():
    k = ModelKernels()
    I, T = torch.randn(8, 16), torch.randn(8, 16)
    # python float scale, no grad anywhere
    l0 = xt.ClipLoss(_kernels=k)(I, T, 7.5)
    l1 = oracle.clip_loss_single(I.double(), T.double(), torch.tensor(7.5, dtype=torch.float64))
    np.testing.assert_allclose(l0.item(), l1.item(), rtol=1e-6)
    # nonscalar_logit_scale: shape [1] scale gets a shape-[1] grad
    s = torch.tensor([7.5], requires_grad=True)
    xt.ClipLoss(_kernels=k)(I, T, s).backward()
    assert s.grad.shape == (1,)
    # upstream gradient is honoured (GradScaler multiplies the loss)
    I2 = I.clone().requires_grad_(True)
    (xt.ClipLoss(_kernels=k)(I2, T, 7.5) * 3.0).backward()
    I3 = I.clone().requires_grad_(True)
    xt.ClipLoss(_kernels=k)(I3, T, 7.5).backward()
    np.testing.assert_allclose(I2.grad.numpy(), 3.0 * I3.grad.numpy(), rtol=1e-5, atol=1e-8)
    # label cache semantics of get_ground_truth (loss.py:91-102)
    m = xt.ClipLoss(local_loss=True, cache_labels=True, rank=2, world_size=4)
    lab = m.get_ground_truth(torch.device("cpu"), 5)
    assert lab.tolist() == [10, 11, 12, 13, 14] and m.labels[torch.device("cpu")] is lab
    assert xt.ClipLoss(rank=2, world_size=4).get_ground_truth(torch.device("cpu"), 3).tolist() == [0, 1, 2]
    # errors
    with pytest.raises(NotImplementedError):
        xt.ClipLoss(use_horovod=True)
    with pytest.raises(ValueError):
        xt.ClipLoss(_kernels=k)(torch.randn(4, 8), torch.randn(5, 8), 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        xt.ClipLoss()(I, T, 7.5)          # product path on CPU tensors must fail loudly


def test_forward_chunks_gradient_accumulation_path():
    """train_other.py:140-197: cached no-grad features + one live chunk.  forward_chunks equals the reference's
    cat-then-loss (loss, live-chunk gradients, d logit_scale) while skipping gradient work for the cached chunks."""
    g = torch.Generator().manual_seed(4)
    sizes, D = [5, 3, 4], 12
    imgs = [torch.nn.functional.normalize(torch.randn(b, D, generator=g, dtype=torch.float64), dim=-1) for b in sizes]
    txts = [torch.nn.functional.normalize(torch.randn(b, D, generator=g, dtype=torch.float64), dim=-1) for b in sizes]
    for live in range(3):
        k = ModelKernels()
        ic = [t.clone().requires_grad_(j == live) for j, t in enumerate(imgs)]
        tc = [t.clone().requires_grad_(j == live) for j, t in enumerate(txts)]
        s = torch.tensor(11.0, dtype=torch.float64, requires_grad=True)
        loss = xt.ClipLoss(_kernels=k).forward_chunks(ic, tc, s)
        loss.backward()
        # reference semantics: autograd through the concatenation
        ir = [t.clone().requires_grad_(j == live) for j, t in enumerate(imgs)]
        tr = [t.clone().requires_grad_(j == live) for j, t in enumerate(txts)]
        sr = torch.tensor(11.0, dtype=torch.float64, requires_grad=True)
        lr = oracle.clip_loss_single(torch.cat(ir), torch.cat(tr), sr)
        lr.backward()
        assert rel_err(loss.item(), lr.item()) < 1e-6
        assert rel_err(ic[live].grad.numpy(), ir[live].grad.numpy()) < 1e-5
        assert rel_err(tc[live].grad.numpy(), tr[live].grad.numpy()) < 1e-5
        assert abs(s.grad.item() - sr.grad.item()) < 1e-5 * abs(sr.grad.item()) + 1e-7
        assert all(t.grad is None for j, t in enumerate(ic) if j != live)
        # gradient work only for the live blocks: one dS-only pass (logit_scale) + two rectangular calls
        bwd = [c for c in k.calls if c[0] == "clip_bwd"]
        lo = sum(sizes[:live])
        assert [c[1][0] for c in bwd] == [sum(sizes), sizes[live], sizes[live]] and all(c[3] in (0, lo) for c in bwd)
    # everything live -> the ordinary fused path; output_dict and validation
    k = ModelKernels()
    ic = [t.clone().requires_grad_(True) for t in imgs]
    tc = [t.clone().requires_grad_(True) for t in txts]
    out = xt.ClipLoss(_kernels=k).forward_chunks(ic, tc, 11.0, output_dict=True)
    assert set(out) == {"contrastive_loss"}
    out["contrastive_loss"].backward()
    assert len([c for c in k.calls if c[0] == "clip_bwd"]) == 1
    with pytest.raises(ValueError):
        xt.ClipLoss(_kernels=k).forward_chunks(ic, tc[:2], 11.0)


@pytest.mark.parametrize("seed", range(12))
def test_forward_chunks_random_layouts(seed):
    """Random chunkings and random subsets of live image / text chunks (not necessarily the same index on both sides):
    value and every live gradient equal autograd through the concatenation; dead chunks receive nothing."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 5))
    sizes = [int(v) for v in rng.integers(1, 7, n)]
    D = int(rng.integers(2, 9))
    live_i = [bool(v) for v in rng.integers(0, 2, n)]
    live_t = [bool(v) for v in rng.integers(0, 2, n)]
    g = torch.Generator().manual_seed(seed)
    imgs = [torch.randn(b, D, generator=g, dtype=torch.float64) for b in sizes]
    txts = [torch.randn(b, D, generator=g, dtype=torch.float64) for b in sizes]
    ic = [t.clone().requires_grad_(l) for t, l in zip(imgs, live_i)]
    tc = [t.clone().requires_grad_(l) for t, l in zip(txts, live_t)]
    ir = [t.clone().requires_grad_(l) for t, l in zip(imgs, live_i)]
    tr = [t.clone().requires_grad_(l) for t, l in zip(txts, live_t)]
    s = torch.tensor(2.5, dtype=torch.float64, requires_grad=True)
    sr = torch.tensor(2.5, dtype=torch.float64, requires_grad=True)
    loss = xt.ClipLoss(_kernels=ModelKernels()).forward_chunks(ic, tc, s)
    lr = oracle.clip_loss_single(torch.cat(ir), torch.cat(tr), sr)
    assert rel_err(loss.item(), lr.item()) < 1e-6
    loss.backward()
    lr.backward()
    for mine, ref in zip(ic + tc, ir + tr):
        if ref.requires_grad:
            # absolute floor: a 1 x 1 problem has an exactly zero gradient in the reference, the fp32 LSEs of the
            # kernel contract leave ~1e-10
            a, b = mine.grad.numpy(), ref.grad.numpy()
            assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max() + 1e-8
        else:
            assert mine.grad is None
    assert abs(s.grad.item() - sr.grad.item()) < 1e-5 * abs(sr.grad.item()) + 1e-7


def test_get_logits_slow_path_matches_reference_expression():
    I, T = torch.randn(6, 8), torch.randn(6, 8)
    li, lt = xt.ClipLoss().get_logits(I, T, torch.tensor(3.0))
    ri, rt = oracle.clip_logits(I, T, torch.tensor(3.0))
    assert torch.equal(li, ri) and torch.equal(lt, rt)


def test_create_loss_contract():
    args = types.SimpleNamespace(distill=False, siglip=False, model="ViT-B-32", local_loss=True,
                                 gather_with_grad=True, rank=3, world_size=8, horovod=False)
    m = xt.create_loss(args)
    assert isinstance(m, xt.ClipLoss)
    assert (m.local_loss, m.gather_with_grad, m.cache_labels, m.rank, m.world_size) == (True, True, True, 3, 8)
    args.siglip, args.loss_dist_impl = True, "gather"
    sg = xt.create_loss(args)                      # factory.py:455-461
    assert isinstance(sg, xt.SigLipLoss) and (sg.rank, sg.world_size, sg.dist_impl) == (3, 8, "gather")
    args.siglip, args.distill = False, True
    with pytest.raises(NotImplementedError):
        xt.create_loss(args)


def test_l2_normalize_golden(golden_dir):
    g = _load(golden_dir, "l2norm.npz")
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    y = xt.l2_normalize(x, _kernels=ModelKernels())
    y.backward(torch.from_numpy(g["gy"]))
    np.testing.assert_allclose(y.detach().numpy(), g["y"], rtol=1e-11, atol=1e-12)
    assert rel_err(x.grad.numpy()[[0, 1, 3, 4, 6, 7, 8]], g["gx"][[0, 1, 3, 4, 6, 7, 8]]) < 1e-6   # inv_norm is fp32 in the ABI
    assert rel_err(x.grad.numpy()[[2, 5]], g["gx"][[2, 5]]) < 1e-6                             # clamped rows: gy / eps


def test_asymmetric_loss_golden(golden_dir):
    g = _load(golden_dir, "asl.npz")
    for n in range(3):
        gn, gp, clip = g[f"k{n}_cfg"]
        x = torch.from_numpy(g["x"]).float().requires_grad_(True)
        m = xt.AsymmetricLoss(gamma_neg=gn, gamma_pos=gp, clip=clip, _kernels=ModelKernels())
        loss = m(x, torch.from_numpy(g["y"]).float())
        loss.backward()
        np.testing.assert_allclose(loss.item(), g[f"k{n}_loss"], rtol=1e-5)
        np.testing.assert_allclose(x.grad.numpy(), g[f"k{n}_dx"], rtol=1e-4, atol=1e-6)
    assert torch.is_grad_enabled()
    d = inspect.signature(xt.AsymmetricLoss.__init__).parameters
    assert (d["gamma_neg"].default, d["gamma_pos"].default, d["clip"].default, d["eps"].default) == (4, 1, 0.05, 1e-8)


def test_tag_head_golden_and_state_dict(golden_dir):
    g = _load(golden_dir, "tag_head.npz")
    for n in range(int(g["n_cases"])):
        pre = f"t{n}_"
        seed, D, b, N, gain = g[pre + "cfg"]
        params = make_tag_params(int(seed), int(D), gain=float(gain), dtype=torch.float64)
        head = xt.TagHead(int(D), tag_list=list(g[pre + "tag_list"]), _kernels=ModelKernels()).double()
        assert sorted(head.state_dict().keys()) == list(g[pre + "state_keys"])   # same 35 keys as the reference
        head.load_state_dict(params, strict=True)
        head.eval()
        tokens = torch.from_numpy(g[pre + "tokens"]).requires_grad_(True)
        logits = head.tag_forward(tokens)
        assert logits.shape == (int(b), 44)
        logits.backward(torch.from_numpy(g[pre + "glogits"]))
        np.testing.assert_allclose(logits.detach().numpy(), g[pre + "logits"], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(tokens.grad.numpy(), g[pre + "dtokens"], rtol=1e-8, atol=1e-11)
        q0 = head.tag_head.encoder.layer[0].crossattention.self.query.weight.grad[:4, :8]
        k1 = head.tag_head.encoder.layer[1].crossattention.self.key.weight.grad[:4, :8]
        np.testing.assert_allclose(q0.numpy(), g[pre + "dq0w"], rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(k1.numpy(), g[pre + "dk1w"], rtol=1e-8, atol=1e-11)
        assert head.prepare_control_words(logits.detach()) == list(g[pre + "words"])
    with pytest.raises(ValueError):
        head.tag_forward(torch.randn(2, 5, int(D) + 1, dtype=torch.float64))


def test_siglip_w1_host_logic_vs_reference_fixture(golden_dir):
    """SigLipLoss (single process) with the contract-model kernels against what the reference's SigLipLoss produced
    (tests/golden/siglip.npz): value, feature / scale / bias gradients, upstream-gradient scaling, output_dict, and the
    no-grad path that skips staging the logit gradient."""
    g = _load(golden_dir, "siglip.npz")
    for n in range(int(g["n_w1"])):
        pre = f"w1_{n}_"
        k = ModelKernels()
        I = torch.from_numpy(g[pre + "I"]).clone().requires_grad_(True)
        T = torch.from_numpy(g[pre + "T"]).clone().requires_grad_(True)
        s = torch.tensor(float(g[pre + "scale"]), dtype=torch.float64, requires_grad=True)
        b = torch.tensor(float(g[pre + "bias"]), dtype=torch.float64, requires_grad=True)
        out = xt.SigLipLoss(compute_dtype=torch.bfloat16, _kernels=k)(I, T, s, b, output_dict=True)
        assert list(out) == ["contrastive_loss"]
        (out["contrastive_loss"] * 2.0).backward()
        # the drop-in rounds the features to bf16 (its compute dtype); the fixture is the reference's fp64 run
        assert rel_err(out["contrastive_loss"].item(), float(g[pre + "loss"])) < 2e-2
        assert rel_err(I.grad.numpy(), 2.0 * g[pre + "dI"]) < 3e-2 and rel_err(T.grad.numpy(), 2.0 * g[pre + "dT"]) < 3e-2
        assert abs(float(s.grad) - 2.0 * float(g[pre + "dscale"])) <= 3e-2 * abs(2.0 * float(g[pre + "dscale"])) + 1e-6
        assert abs(float(b.grad) - 2.0 * float(g[pre + "dbias"])) <= 3e-2 * abs(2.0 * float(g[pre + "dbias"])) + 1e-6
        assert [c[0] for c in k.calls] == ["siglip_fwd", "clip_bwd"]
    with torch.no_grad():
        k = ModelKernels()
        xt.SigLipLoss(compute_dtype=torch.bfloat16, _kernels=k)(I.detach(), T.detach(), 10.0, -10.0)
        assert [c[0] for c in k.calls] == ["siglip_fwd"]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        xt.SigLipLoss()(I.detach(), T.detach(), 10.0, -10.0)
    with pytest.raises(NotImplementedError):
        xt.SigLipLoss(_kernels=ModelKernels())(I.detach().float(), T.detach().float(), 10.0, -10.0)   # fp32, no autocast


def test_tag_head_fused_paths_host_logic():
    """The production plumbing of the tag head (one fused K|V projection for both layers, K4 reading / writing column
    slices of shared buffers, K6 dense-output blocks, projection backward on the package's GEMMs) with the contract
    model on CPU, bf16, against the library path of the same head (one Linear per projection, torch LayerNorm):
    logits and the gradients of tokens, projection weights / biases, LayerNorm parameters and the label embeddings."""
    torch.manual_seed(0)
    D, b, N = 64, 3, 7
    params = make_tag_params(3, D, gain=4.0, dtype=torch.float32)
    g = torch.Generator().manual_seed(1)
    tokens = torch.randn(b, N, D, generator=g)
    wgt = torch.randn(b, 44, generator=g)
    outs, calls = [], []
    for fused in (True, False):
        k = ModelKernels()
        head = xt.TagHead(D, fuse_kv=fused, fuse_ln=fused, _kernels=k)
        head.load_state_dict(params, strict=True)
        head = head.bfloat16().eval()
        tok = tokens.bfloat16().requires_grad_(True)
        logits = head.tag_forward(tok)
        (logits.float() * wgt).sum().backward()
        L = head.tag_head.encoder.layer
        outs.append([logits.float().detach(), tok.grad.float(),
                     L[1].crossattention.self.key.weight.grad.float(), L[0].crossattention.self.value.bias.grad.float(),
                     L[0].crossattention.output.LayerNorm.weight.grad.float(), L[1].output.LayerNorm.bias.grad.float(),
                     head.tag_labels.weight.grad.float()])
        calls.append([c[0] for c in k.calls])
    for a, c in zip(outs[0], outs[1]):
        assert rel_err(a.numpy(), c.numpy()) < 6e-2, rel_err(a.numpy(), c.numpy())
    fused_calls = calls[0]
    assert fused_calls.count("tc_linear") == 1 and fused_calls.count("tc_gemm") == 2      # projection fwd; dX and dW
    assert fused_calls.count("ln_res_fwd") == 4 and fused_calls.count("ln_res_bwd") == 4
    assert "tc_linear" not in calls[1] and "ln_res_fwd" not in calls[1]
