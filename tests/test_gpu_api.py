"""GPU parity tests, API level: the reference-facing Python API (ClipLoss / TagHead / AsymmetricLoss / l2_normalize)
running on the CUDA library, against the golden fixtures (reference outputs) and the oracle; plus full-size
(BASELINE config) property / known-answer tests where the oracle would take minutes."""
import math
import os

import numpy as np
import pytest
import torch

import oracle
import xtag_clip_b200 as xt
from oracle.tag_oracle import make_tag_params

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def feats(seed, b, d, corr=0.3):
    g = torch.Generator().manual_seed(seed)
    i = torch.randn(b, d, generator=g)
    n = torch.randn(b, d, generator=g)
    t = corr * i + (1 - corr) * n
    return torch.nn.functional.normalize(i, dim=-1), torch.nn.functional.normalize(t, dim=-1)


def test_cliploss_fp32_mode_golden(golden_dir):
    """fp32 inputs -> exact SIMT path; bar 1e-5 against the reference's own outputs."""
    g = np.load(os.path.join(golden_dir, "clip_w1.npz"))
    for n in range(int(g["n_cases"])):
        pre = f"c{n}_f64_"
        I = torch.from_numpy(g[pre + "I"]).float().cuda().requires_grad_(True)
        T = torch.from_numpy(g[pre + "T"]).float().cuda().requires_grad_(True)
        s = torch.tensor(float(g[pre + "scale"]), device="cuda", requires_grad=True)
        loss = xt.ClipLoss()(I, T, s)
        assert loss.dtype == torch.float32 and loss.dim() == 0 and loss.is_cuda
        loss.backward()
        assert rel_err(loss, g[pre + "loss"]) < 1e-5
        assert rel_err(I.grad, g[pre + "dI"]) < 1e-5
        assert rel_err(T.grad, g[pre + "dT"]) < 1e-5
        assert abs(s.grad.item() - float(g[pre + "dscale"])) < 1e-3 * abs(float(g[pre + "dscale"])) + 2e-5


@pytest.mark.parametrize("B,D,scale", [(256, 512, 14.285714), (1024, 512, 14.285714), (1024, 512, 100.0),
                                       (2048, 512, 14.285714), (4096, 512, 14.285714), (1000, 768, 30.0)])
def test_cliploss_bf16_config2_vs_oracle(B, D, scale):
    """BASELINE config 2 (ViT-B-32 head sweep, D=512, bf16): loss <= 1e-3, grads <= 2e-2 vs the oracle evaluated in
    fp64 on the SAME bf16-rounded inputs."""
    I, T = feats(B + D, B, D, corr=0.15 if scale > 50 else 0.4)
    Ib, Tb = I.bfloat16(), T.bfloat16()
    Ic = Ib.cuda().requires_grad_(True)
    Tc = Tb.cuda().requires_grad_(True)
    s = torch.tensor(scale, device="cuda", requires_grad=True)
    out = xt.ClipLoss()(Ic, Tc, s, output_dict=True)
    out["contrastive_loss"].backward()
    assert Ic.grad.dtype == torch.bfloat16
    torch.set_num_threads(os.cpu_count() or 8)
    lo, dI, dT, ds, _, _ = oracle.clip_loss_closed_form(Ib.double(), Tb.double(), torch.tensor(scale, dtype=torch.float64))
    assert rel_err(out["contrastive_loss"], lo) < 1e-3
    assert rel_err(Ic.grad, dI) < 2e-2
    assert rel_err(Tc.grad, dT) < 2e-2
    assert abs(s.grad.item() - ds.item()) < 2e-2 * abs(ds.item()) + 2e-5


def test_cliploss_fp16_and_mixed_inputs_upcast():
    I, T = feats(3, 64, 32)
    a = xt.ClipLoss()(I.half().cuda(), T.half().cuda(), 10.0)
    b = oracle.clip_loss_single(I.half().double(), T.half().double(), torch.tensor(10.0, dtype=torch.float64))
    assert rel_err(a, b) < 1e-5
    c = xt.ClipLoss()(I.bfloat16().cuda(), T.cuda(), 10.0)       # mixed -> fp32
    d = oracle.clip_loss_single(I.bfloat16().double(), T.double(), torch.tensor(10.0, dtype=torch.float64))
    assert rel_err(c, d) < 1e-5


def test_cliploss_full_size_known_answer():
    """BASELINE config 5 shape on one GPU (B=32768, D=1024, bf16): one-hot features I_i = T_i = e_{i mod D} have a
    closed-form loss and gradient: S_ij = s if i == j (mod D) else 0."""
    B, D, s = 32768, 1024, 5.0
    idx = torch.arange(B, device="cuda") % D
    I = torch.zeros(B, D, device="cuda", dtype=torch.bfloat16)
    I[torch.arange(B, device="cuda"), idx] = 1.0
    T = I.clone()
    I.requires_grad_(True)
    T.requires_grad_(True)
    sc = torch.tensor(s, device="cuda", requires_grad=True)
    loss = xt.ClipLoss()(I, T, sc)
    loss.backward()
    r = B // D
    Z = r * math.exp(s) + (B - r)
    expect = math.log(Z) - s
    assert abs(loss.item() - expect) < 1e-3 * expect
    # dS_ij = (P_row + P_col - 2 delta)/(2B) with P_row = P_col = e^{S_ij}/Z  =>  dI_i = s * [ (2 r e^s/Z - 2)/(2B) e_c(i)
    #                                                                        + sum_{c != c(i)} (2 r / Z)/(2B) e_c ]
    on = s * (2 * r * math.exp(s) / Z - 2) / (2 * B)
    off = s * (2 * r / Z) / (2 * B)
    gi = I.grad.float()
    assert rel_err(gi[torch.arange(B, device="cuda"), idx], torch.full((B,), on)) < 2e-2
    sel = torch.ones(B, D, dtype=torch.bool, device="cuda")
    sel[torch.arange(B, device="cuda"), idx] = False
    assert abs(gi[sel].mean().item() - off) < 2e-2 * off and abs(gi[sel].max().item() - off) < 5e-2 * off
    assert rel_err(T.grad.float(), gi) < 1e-2                       # symmetric problem
    ds = ((2 * r * math.exp(s) / Z - 2) / (2 * B)) * B               # sum_ij dS_ij S_ij / s: only S_ij = s entries count
    # off-diagonal (mod D) entries have S = 0; per row r entries at s: r*(2 e^s/Z)/(2B) - 2/(2B)
    assert abs(sc.grad.item() - ds) < 2e-2 * abs(ds)


def test_cliploss_full_size_properties():
    """Size-independent identities at B=32768, D=1024 (random bf16 features):
       <I, dI> = <T, dT> = s * d(logit_scale)   (dI = s dS T, dT = s dS^T I, ds = sum dS.S / s)
       swapping the roles of image and text leaves the loss unchanged and swaps the gradients."""
    B, D, s = 32768, 1024, 14.285714
    I, T = feats(77, B, D, corr=0.3)
    I, T = I.bfloat16().cuda(), T.bfloat16().cuda()
    Ia, Ta = I.clone().requires_grad_(True), T.clone().requires_grad_(True)
    sa = torch.tensor(s, device="cuda", requires_grad=True)
    la = xt.ClipLoss()(Ia, Ta, sa)
    la.backward()
    ii = (Ia.grad.float() * I.float()).sum().item()
    tt = (Ta.grad.float() * T.float()).sum().item()
    assert abs(ii - tt) < 2e-2 * abs(ii) + 1e-5
    assert abs(ii - s * sa.grad.item()) < 2e-2 * abs(ii) + 1e-5
    Ib, Tb = T.clone().requires_grad_(True), I.clone().requires_grad_(True)
    lb = xt.ClipLoss()(Ib, Tb, torch.tensor(s, device="cuda"))
    lb.backward()
    assert abs(la.item() - lb.item()) < 1e-4 * la.item()
    assert rel_err(Ib.grad, Ta.grad) < 2e-2 and rel_err(Tb.grad, Ia.grad) < 2e-2
    assert 0.0 < la.item() < math.log(B)


def test_tag_head_golden_fp32(golden_dir):
    g = np.load(os.path.join(golden_dir, "tag_head.npz"))
    for n in range(int(g["n_cases"])):
        pre = f"t{n}_"
        seed, D, b, N, gain = g[pre + "cfg"]
        params = make_tag_params(int(seed), int(D), gain=float(gain), dtype=torch.float32)
        head = xt.TagHead(int(D), tag_list=list(g[pre + "tag_list"])).cuda()
        head.load_state_dict(params, strict=True)
        head.eval()
        tokens = torch.from_numpy(g[pre + "tokens"]).float().cuda().requires_grad_(True)
        torch.backends.cuda.matmul.allow_tf32 = False      # the dense layers are torch library calls
        logits = head.tag_forward(tokens)
        logits.backward(torch.from_numpy(g[pre + "glogits"]).float().cuda())
        assert rel_err(logits, g[pre + "logits"]) < 2e-5
        assert rel_err(tokens.grad, g[pre + "dtokens"]) < 1e-4
        assert head.prepare_control_words(logits.detach()) == list(g[pre + "words"])


def test_tag_head_bf16_config3_shape():
    """BASELINE config 3 shape (tokens [b,197,512], b reduced to keep the fp64 oracle in seconds), bf16 autocast
    as the reference trains (precision amp_bf16)."""
    D, b, N = 512, 32, 197
    params = make_tag_params(7, D, gain=4.0, dtype=torch.float32)
    head = xt.TagHead(D).cuda()
    head.load_state_dict(params, strict=True)
    head.eval()
    g = torch.Generator().manual_seed(8)
    tokens = torch.randn(b, N, D, generator=g)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = head.tag_forward(tokens.cuda())
    p64 = {k: v.double() for k, v in params.items()}
    ref = oracle.tag_head_forward(tokens.double(), p64)
    assert rel_err(logits.float(), ref) < 3e-2


def test_tag_head_bf16_config3_full_batch_fwd_bwd():
    """BASELINE config 3 at its real size (tokens [1024, 197, 512], bf16 autocast, eval): forward AND token / weight
    gradients of the production path (fused K|V projection GEMM + K4) against the fp64 oracle on a 32-sample slice
    (samples are independent; bars: logits 3e-2, gradients 5e-2 max-norm), and against the unfused path (one library
    Linear per projection) on the full batch."""
    D, b, N, ns = 512, 1024, 197, 32
    params = make_tag_params(11, D, gain=4.0, dtype=torch.float32)
    heads = [xt.TagHead(D).cuda(), xt.TagHead(D, fuse_kv=False, fuse_ln=False).cuda()]     # production / library path
    for h in heads:
        h.load_state_dict(params, strict=True)
        h.eval()
    g = torch.Generator().manual_seed(12)
    tokens = torch.randn(b, N, D, generator=g)
    wgt = torch.randn(b, 44, generator=g)
    outs = []
    for h in heads:
        tok = tokens.cuda().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = h.tag_forward(tok)
        (logits.float() * wgt.cuda()).sum().backward()
        kw = h.tag_head.encoder.layer[1].crossattention.self.key.weight.grad
        vb = h.tag_head.encoder.layer[0].crossattention.self.value.bias.grad
        lw = h.tag_head.encoder.layer[0].crossattention.output.LayerNorm.weight.grad
        lb = h.tag_head.encoder.layer[1].output.LayerNorm.bias.grad
        outs.append((logits.float().detach(), tok.grad.detach(), kw.detach().clone(), vb.detach().clone(),
                     lw.detach().clone(), lb.detach().clone(), h.tag_labels.weight.grad.detach().clone()))
    for a, c in zip(outs[0], outs[1]):                                   # two bf16 pipelines, max-norm
        assert rel_err(a, c) < 4e-2, rel_err(a, c)
    p64 = {k: v.double() for k, v in params.items()}
    tok64 = tokens[:ns].double().requires_grad_(True)
    ref = oracle.tag_head_forward(tok64, p64)
    (ref * wgt[:ns].double()).sum().backward()
    assert rel_err(outs[0][0][:ns], ref) < 3e-2
    assert rel_err(outs[0][1][:ns], tok64.grad) < 5e-2


def test_asymmetric_loss_and_l2_api(golden_dir):
    g = np.load(os.path.join(golden_dir, "asl.npz"))
    x = torch.from_numpy(g["x"]).float().cuda().requires_grad_(True)
    loss = xt.AsymmetricLoss()(x, torch.from_numpy(g["y"]).float())
    (2.0 * loss).backward()
    assert rel_err(loss, g["k0_loss"]) < 1e-5 and rel_err(x.grad, 2.0 * g["k0_dx"]) < 1e-5
    gl = np.load(os.path.join(golden_dir, "l2norm.npz"))
    xx = torch.from_numpy(gl["x"]).float().cuda().requires_grad_(True)
    y = xt.l2_normalize(xx)
    y.backward(torch.from_numpy(gl["gy"]).float().cuda())
    assert rel_err(y, gl["y"]) < 1e-6
    rows = [0, 1, 3, 4, 6, 7, 8]
    assert rel_err(xx.grad[rows], gl["gx"][rows]) < 1e-5


def test_launch_counter_counts_native_kernels():
    from xtag_clip_b200 import _lib
    n0 = _lib.launch_count()
    I, T = feats(1, 256, 512)
    xt.ClipLoss()(I.bfloat16().cuda(), T.bfloat16().cuda(), 10.0)
    assert _lib.launch_count() - n0 >= 3          # tcgen05 fwd + fused row/column reduction + loss


def test_cliploss_cuda_graph_replay_matches_eager():
    """cuda_graph=True captures forward and backward once and replays them: same numbers as the eager path, for
    fresh inputs on every call (inputs are copied into the captured step's static buffers)."""
    from xtag_clip_b200 import _lib
    graphed, eager = xt.ClipLoss(cuda_graph=True), xt.ClipLoss()
    counts = []
    for it in range(4):
        counts.append(_lib.launch_count())
        I, T = feats(100 + it, 512, 256, corr=0.3)
        out = []
        for mod in (graphed, eager):
            Ic = I.bfloat16().cuda().requires_grad_(True)
            Tc = T.bfloat16().cuda().requires_grad_(True)
            ls = torch.tensor(2.0 + 0.2 * it, device="cuda", requires_grad=True)
            loss = mod(Ic, Tc, ls.exp())
            (loss * (1.0 + it)).backward()
            out.append((loss.detach().clone(), Ic.grad.clone(), Tc.grad.clone(), ls.grad.clone()))
        for a, b in zip(out[0], out[1]):
            assert rel_err(a, b) < 1e-6
    assert len(graphed._graphs) == 1
    # replays are credited to the library's launch counter (bench.py's gpu_launches): every iteration launches the
    # same number of library kernels through the graph as through the eager module
    per_iter = [b - a for a, b in zip(counts, counts[1:])]
    assert per_iter[1] == per_iter[2] > 10 and per_iter[1] % 2 == 0


def test_config1_heads_on_reference_encoder_outputs(golden_dir):
    """BASELINE config 1: the hot path on the outputs of the reference's own ViT-B-32 encoders (batch 16, CPU fp32 run
    of the unmodified reference, tests/golden/config1.npz): ClipLoss, tag head + AsymmetricLoss and their gradients."""
    g = np.load(os.path.join(golden_dir, "config1.npz"))
    I = torch.from_numpy(g["image_features"]).cuda().requires_grad_(True)
    T = torch.from_numpy(g["text_features"]).cuda().requires_grad_(True)
    s = torch.tensor(float(g["logit_scale"]), device="cuda", requires_grad=True)
    tok = torch.from_numpy(g["tokens"]).cuda().requires_grad_(True)
    head = xt.TagHead(512, tag_list=list(g["tag_list"])).cuda()
    head.load_state_dict(make_tag_params(50, 512, gain=4.0, dtype=torch.float32), strict=True)
    head.eval()
    torch.backends.cuda.matmul.allow_tf32 = False
    logits = head.tag_forward(tok)
    closs = xt.ClipLoss()(I, T, s, output_dict=True)["contrastive_loss"]
    tloss = xt.AsymmetricLoss()(logits, torch.from_numpy(g["additional"]).repeat(1, 2))
    (closs + tloss).backward()
    assert rel_err(closs, g["contrastive_loss"]) < 1e-5
    assert rel_err(I.grad, g["d_image_features"]) < 2e-5 and rel_err(T.grad, g["d_text_features"]) < 2e-5
    assert rel_err(logits, g["tag_logits"]) < 1e-4
    assert rel_err(tloss, g["tag_loss"]) < 1e-4
    assert rel_err(tok.grad[:2, :4, :16], g["d_tokens_head"]) < 1e-3
    assert abs(tok.grad.norm().item() - float(g["d_tokens_norm"])) < 1e-3 * float(g["d_tokens_norm"])
    assert head.prepare_control_words(logits.detach()) == list(g["words"])
    # the same step under bf16 autocast, as the reference trains (--precision amp_bf16): BASELINE bf16 bars
    I2 = torch.from_numpy(g["image_features"]).cuda().bfloat16().requires_grad_(True)
    T2 = torch.from_numpy(g["text_features"]).cuda().bfloat16().requires_grad_(True)
    l2 = xt.ClipLoss()(I2, T2, torch.tensor(float(g["logit_scale"]), device="cuda"))
    l2.backward()
    ref = oracle.clip_loss_closed_form(I2.detach().double().cpu(), T2.detach().double().cpu(),
                                       torch.tensor(float(g["logit_scale"]), dtype=torch.float64))
    assert rel_err(l2, ref[0]) < 1e-3 and rel_err(I2.grad, ref[1]) < 2e-2 and rel_err(T2.grad, ref[2]) < 2e-2


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_forward_chunks_gradient_accumulation_gpu(dtype, tol):
    """train_other.py:140-197 on the real kernels: cached no-grad chunks + one live chunk against the oracle."""
    g = torch.Generator().manual_seed(8)
    sizes, D = [256, 128, 384], 256
    nrm = torch.nn.functional.normalize
    imgs = [nrm(torch.randn(b, D, generator=g), dim=-1).to(dtype) for b in sizes]
    txts = [nrm(0.3 * i.float() + 0.7 * torch.randn(i.shape[0], D, generator=g), dim=-1).to(dtype) for i in imgs]
    live = 1
    ic = [t.cuda().requires_grad_(j == live) for j, t in enumerate(imgs)]
    tc = [t.cuda().requires_grad_(j == live) for j, t in enumerate(txts)]
    s = torch.tensor(14.285714, device="cuda", requires_grad=True)
    loss = xt.ClipLoss().forward_chunks(ic, tc, s)
    loss.backward()
    I = torch.cat(imgs).double()
    T = torch.cat(txts).double()
    lo, dI, dT, ds, _, _ = oracle.clip_loss_closed_form(I, T, torch.tensor(14.285714, dtype=torch.float64))
    a, b = sum(sizes[:live]), sum(sizes[:live + 1])
    assert rel_err(loss, lo) < (1e-5 if dtype == torch.float32 else 1e-3)
    assert rel_err(ic[live].grad, dI[a:b]) < tol and rel_err(tc[live].grad, dT[a:b]) < tol
    assert abs(float(s.grad) - float(ds)) <= 1e-3 * abs(float(ds)) + 2e-5


def test_l2_normalize_into_graph_input_slots():
    """Features normalised by K3 straight into the captured step's input slots (no input copy at replay): same loss
    and same gradients w.r.t. the RAW features as normalise -> eager loss."""
    B, D = 512, 256
    mod = xt.ClipLoss(cuda_graph=True)
    g = torch.Generator().manual_seed(4)
    raw = [(torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)) for _ in range(3)]
    s = torch.tensor(14.285714, device="cuda")
    # first call captures the step; from then on the slots exist
    I0 = xt.l2_normalize(raw[0][0].cuda(), out_dtype=torch.bfloat16).detach().requires_grad_(True)
    T0 = xt.l2_normalize(raw[0][1].cuda(), out_dtype=torch.bfloat16).detach().requires_grad_(True)
    mod(I0, T0, s).backward()
    slots = mod.graph_input_slots(B, D, torch.bfloat16)
    assert slots is not None
    for ri, rt in raw[1:]:
        xi, xt_ = ri.cuda().requires_grad_(True), rt.cuda().requires_grad_(True)
        I = xt.l2_normalize(xi, out=slots[0])
        T = xt.l2_normalize(xt_, out=slots[1])
        assert I.data_ptr() == slots[0].data_ptr() and I.requires_grad
        loss = mod(I, T, s)
        loss.backward()
        yi, yt = ri.cuda().requires_grad_(True), rt.cuda().requires_grad_(True)
        ref = xt.ClipLoss()(xt.l2_normalize(yi, out_dtype=torch.bfloat16), xt.l2_normalize(yt, out_dtype=torch.bfloat16), s)
        ref.backward()
        assert rel_err(loss, ref) < 1e-6 and rel_err(xi.grad, yi.grad) < 1e-5 and rel_err(xt_.grad, yt.grad) < 1e-5


@pytest.mark.parametrize("B,D,scale,bias", [(256, 512, 10.0, -10.0), (520, 264, 25.0, -4.0), (1024, 1024, 100.0, 0.0),
                                            (130, 72, 1.0, 0.5)])
def test_siglip_loss_vs_oracle(B, D, scale, bias):
    """Fused SigLipLoss (one pass: loss + staged logit gradient; backward = two GEMMs) against the oracle restatement
    of the reference (loss.py:314-448) on the same bf16 features: loss 1e-3, gradients 2e-2 (max-norm)."""
    I, T = feats(B + D, B, D, corr=0.4)
    Ib, Tb = I.bfloat16(), T.bfloat16()
    Ic, Tc = Ib.cuda().requires_grad_(True), Tb.cuda().requires_grad_(True)
    s = torch.tensor(scale, device="cuda", requires_grad=True)
    b = torch.tensor(bias, device="cuda", requires_grad=True)
    loss = xt.SigLipLoss()(Ic, Tc, s, b)
    (loss * 1.5).backward()
    lo, dI, dT, ds, db = oracle.siglip_loss_world([Ib.double()], [Tb.double()], scale, bias)
    assert rel_err(loss, lo[0]) < 1e-3
    assert rel_err(Ic.grad, 1.5 * dI[0]) < 2e-2 and rel_err(Tc.grad, 1.5 * dT[0]) < 2e-2
    assert abs(float(s.grad) - 1.5 * float(ds[0])) <= 2e-2 * abs(1.5 * float(ds[0])) + 1e-4
    assert abs(float(b.grad) - 1.5 * float(db[0])) <= 2e-2 * abs(1.5 * float(db[0])) + 1e-4
    with torch.no_grad():                                   # evaluation: no gradient staging
        assert rel_err(xt.SigLipLoss()(Ic, Tc, s, b), lo[0]) < 1e-3


def test_siglip_golden_fp32_inputs_under_autocast(golden_dir):
    """Reference-generated fixture (tests/golden/siglip.npz) through the drop-in under bf16 autocast."""
    g = np.load(os.path.join(golden_dir, "siglip.npz"))
    pre = "w1_1_"
    D = g[pre + "I"].shape[1]
    I = torch.from_numpy(g[pre + "I"]).float().cuda()
    T = torch.from_numpy(g[pre + "T"]).float().cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = xt.SigLipLoss()(I, T, torch.tensor(float(g[pre + "scale"]), device="cuda"),
                               torch.tensor(float(g[pre + "bias"]), device="cuda"))
    assert D % 8 == 0 and rel_err(loss, g[pre + "loss"]) < 2e-2


def test_cliploss_autocast_fp32_features_take_the_bf16_path():
    """The reference's amp_bf16 loop calls the loss inside torch.autocast on fp32 features (F.normalize is promoted to
    fp32); its own matmul then runs in bf16.  The drop-in must do the same: tcgen05 path, identical numbers to passing
    bf16-cast features, gradients returned in fp32; outside autocast fp32 features keep the exact fp32 path."""
    from xtag_clip_b200 import _lib
    I, T = feats(21, 512, 256)
    Ic, Tc = I.cuda().requires_grad_(True), T.cuda().requires_grad_(True)
    s = torch.tensor(14.285714, device="cuda", requires_grad=True)
    lib = _lib.load()
    lib.xtag_prof_enable(1)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = xt.ClipLoss()(Ic, Tc, s)
    loss.backward()
    torch.cuda.synchronize()
    import ctypes
    tags, tms, work = (ctypes.c_int * 64)(), (ctypes.c_float * 64)(), (ctypes.c_double * 64)()
    n = lib.xtag_prof_read(tags, tms, work, 64)
    lib.xtag_prof_enable(0)
    assert n >= 4 and {tags[i] for i in range(n)} == {0, 1, 2}          # the tcgen05 kernels ran (fwd, dS, 2 GEMMs)
    Ib, Tb = I.bfloat16().cuda().requires_grad_(True), T.bfloat16().cuda().requires_grad_(True)
    s2 = torch.tensor(14.285714, device="cuda", requires_grad=True)
    ref = xt.ClipLoss()(Ib, Tb, s2)
    ref.backward()
    assert Ic.grad.dtype == torch.float32 and float(loss) == float(ref)
    assert torch.equal(Ic.grad, Ib.grad.float()) and torch.equal(Tc.grad, Tb.grad.float())
    exact = xt.ClipLoss()(I.cuda(), T.cuda(), 14.285714)                   # no autocast: exact fp32 mode
    lo = oracle.clip_loss_single(I.double(), T.double(), torch.tensor(14.285714, dtype=torch.float64))
    assert rel_err(exact, lo) < 1e-5
    forced = xt.ClipLoss(compute_dtype=torch.bfloat16)(I.cuda(), T.cuda(), 14.285714)
    assert float(forced) == float(ref)
    with pytest.raises(ValueError):
        xt.ClipLoss(compute_dtype=torch.float16)
    assert rel_err(xt.ClipLoss()(I.double().cuda(), T.double().cuda(), 14.285714), lo) < 1e-5    # fp64 in -> fp32 mode


def test_cliploss_cuda_graph_second_forward_before_backward():
    """A captured step has ONE set of saved activations: a second forward of the same signature before the first one's
    backward must not overwrite them (it runs eagerly); both backwards give the right gradients."""
    mod, eager = xt.ClipLoss(cuda_graph=True), xt.ClipLoss()
    pairs = []
    for seed in (31, 32):
        I, T = feats(seed, 384, 256, corr=0.2)
        pairs.append((I.bfloat16().cuda(), T.bfloat16().cuda()))
    warm = [t.clone().requires_grad_(True) for t in pairs[0]]
    mod(warm[0], warm[1], 10.0).backward()                                  # capture
    xs = [[t.clone().requires_grad_(True) for t in p] for p in pairs]
    l0 = mod(xs[0][0], xs[0][1], 10.0)
    l1 = mod(xs[1][0], xs[1][1], 10.0)                                      # first backward still outstanding
    l1.backward()
    l0.backward()
    for (a, b), p in zip(xs, pairs):
        ra, rb = p[0].clone().requires_grad_(True), p[1].clone().requires_grad_(True)
        eager(ra, rb, 10.0).backward()
        assert rel_err(a.grad, ra.grad) < 1e-5 and rel_err(b.grad, rb.grad) < 1e-5
    assert rel_err(l0, eager(pairs[0][0], pairs[0][1], 10.0)) < 1e-6
