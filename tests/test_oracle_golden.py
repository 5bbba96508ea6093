"""Pins the oracle (oracle/) against the golden fixtures that oracle/make_golden.py produced by
executing the unmodified reference (/root/reference) on CPU.  CPU-only; runs everywhere."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle.tag_oracle import make_tag_params

T64 = dict(rtol=1e-11, atol=1e-12)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_l2norm_golden(golden_dir):
    g = _load(golden_dir, "l2norm.npz")
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    y = oracle.l2_normalize(x)
    y.backward(torch.from_numpy(g["gy"]))
    np.testing.assert_allclose(y.detach().numpy(), g["y"], **T64)
    np.testing.assert_allclose(x.grad.numpy(), g["gx"], rtol=1e-10, atol=1e-10)


def test_asl_golden(golden_dir):
    g = _load(golden_dir, "asl.npz")
    for n in range(3):
        gn, gp, clip = g[f"k{n}_cfg"]
        x = torch.from_numpy(g["x"]).requires_grad_(True)
        loss = oracle.asymmetric_loss(x, torch.from_numpy(g["y"]), gn, gp, clip)
        loss.backward()
        np.testing.assert_allclose(loss.detach().numpy(), g[f"k{n}_loss"], **T64)
        np.testing.assert_allclose(x.grad.numpy(), g[f"k{n}_dx"], **T64)


@pytest.mark.parametrize("tag,tol", [("f64", T64), ("f32", dict(rtol=3e-5, atol=2e-6))])
def test_clip_w1_golden(golden_dir, tag, tol):
    g = _load(golden_dir, "clip_w1.npz")
    for n in range(int(g["n_cases"])):
        pre = f"c{n}_{tag}_"
        # the oracle always evaluates in fp64; the f32 fixtures are the reference's fp32 run
        I = torch.from_numpy(g[pre + "I"]).double().requires_grad_(True)
        T = torch.from_numpy(g[pre + "T"]).double().requires_grad_(True)
        s = torch.tensor(float(g[pre + "scale"]), dtype=torch.float64, requires_grad=True)
        loss = oracle.clip_loss_single(I, T, s)
        loss.backward()
        np.testing.assert_allclose(loss.detach().numpy(), g[pre + "loss"], **tol)
        np.testing.assert_allclose(I.grad.numpy(), g[pre + "dI"], **tol)
        np.testing.assert_allclose(T.grad.numpy(), g[pre + "dT"], **tol)
        np.testing.assert_allclose(s.grad.numpy(), g[pre + "dscale"], **tol)
        # closed form == autograd == reference
        cl, cdI, cdT, cds, _, _ = oracle.clip_loss_closed_form(I.detach(), T.detach(), s.detach())
        np.testing.assert_allclose(cl.numpy(), g[pre + "loss"], **tol)
        np.testing.assert_allclose(cdI.numpy(), g[pre + "dI"], **tol)
        np.testing.assert_allclose(cdT.numpy(), g[pre + "dT"], **tol)
        np.testing.assert_allclose(cds.numpy(), g[pre + "dscale"], **tol)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("local_loss", [False, True])
@pytest.mark.parametrize("gwg", [False, True])
def test_clip_dist_golden(golden_dir, world, local_loss, gwg):
    """The single-process emulation of W ranks reproduces what the reference produced on each
    rank under gloo (all four local_loss x gather_with_grad modes)."""
    g = _load(golden_dir, "clip_dist.npz")
    pre = f"w{world}_"
    b = int(g[pre + "b"])
    I = torch.from_numpy(g[pre + "I"])
    T = torch.from_numpy(g[pre + "T"])
    Il = [I[r * b:(r + 1) * b] for r in range(world)]
    Tl = [T[r * b:(r + 1) * b] for r in range(world)]
    losses, dI, dT, ds = oracle.clip_loss_world(Il, Tl, float(g[pre + "scale"]), local_loss, gwg)
    for r in range(world):
        key = pre + f"ll{int(local_loss)}_gwg{int(gwg)}_r{r}_"
        np.testing.assert_allclose(losses[r].numpy(), g[key + "loss"], **T64)
        np.testing.assert_allclose(dI[r].numpy(), g[key + "dI"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(dT[r].numpy(), g[key + "dT"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(ds[r].numpy(), g[key + "dscale"], rtol=1e-10, atol=1e-13)


def test_clip_dist_invariants(golden_dir):
    """mean_r(loss_r) == single-process loss; gather_with_grad grads == W x single."""
    g = _load(golden_dir, "clip_dist.npz")
    for world in (2, 3):
        pre = f"w{world}_"
        b = int(g[pre + "b"])
        I = torch.from_numpy(g[pre + "I"]).requires_grad_(True)
        T = torch.from_numpy(g[pre + "T"]).requires_grad_(True)
        s = torch.tensor(float(g[pre + "scale"]), dtype=torch.float64, requires_grad=True)
        loss = oracle.clip_loss_single(I, T, s)
        loss.backward()
        mean_loss = np.mean([g[pre + f"ll1_gwg1_r{r}_loss"] for r in range(world)])
        np.testing.assert_allclose(mean_loss, loss.item(), rtol=1e-9)
        for r in range(world):
            np.testing.assert_allclose(g[pre + f"ll1_gwg1_r{r}_dI"], world * I.grad[r * b:(r + 1) * b].numpy(),
                                       rtol=1e-9, atol=1e-13)
            np.testing.assert_allclose(g[pre + f"ll0_gwg0_r{r}_dT"], T.grad[r * b:(r + 1) * b].numpy(),
                                       rtol=1e-9, atol=1e-13)


def test_tag_head_golden(golden_dir):
    g = _load(golden_dir, "tag_head.npz")
    for n in range(int(g["n_cases"])):
        pre = f"t{n}_"
        seed, D, b, N, gain = g[pre + "cfg"]
        params = make_tag_params(int(seed), int(D), gain=float(gain), dtype=torch.float64)
        assert sorted(params.keys()) == list(g[pre + "state_keys"])      # 35 tag_* keys, SURVEY §5
        for v in params.values():
            v.requires_grad_(True)
        tokens = torch.from_numpy(g[pre + "tokens"]).requires_grad_(True)
        logits = oracle.tag_head_forward(tokens, params)
        logits.backward(torch.from_numpy(g[pre + "glogits"]))
        np.testing.assert_allclose(logits.detach().numpy(), g[pre + "logits"], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(tokens.grad.numpy(), g[pre + "dtokens"], rtol=1e-9, atol=1e-12)
        k0 = "tag_head.encoder.layer.0.crossattention.self.query.weight"
        k1 = "tag_head.encoder.layer.1.crossattention.self.key.weight"
        np.testing.assert_allclose(params[k0].grad[:4, :8].numpy(), g[pre + "dq0w"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(params[k1].grad[:4, :8].numpy(), g[pre + "dk1w"], rtol=1e-9, atol=1e-12)
        idx = oracle.control_word_indices(logits.detach())
        tag_list = list(g[pre + "tag_list"])
        words = [",".join(tag_list[i] for i in row) for row in idx.tolist()]
        assert words == list(g[pre + "words"])


def test_config1_golden(golden_dir):
    """BASELINE config 1 (reference ViT-B-32 XTag model, batch 16, CPU fp32): the oracle reproduces the reference's
    head outputs from the reference's encoder outputs."""
    g = _load(golden_dir, "config1.npz")
    assert int(g["n_params"]) > 150_000_000                      # the full 178.5 M-parameter model produced this
    I = torch.from_numpy(g["image_features"]).double().requires_grad_(True)
    T = torch.from_numpy(g["text_features"]).double().requires_grad_(True)
    s = torch.tensor(float(g["logit_scale"]), dtype=torch.float64, requires_grad=True)
    tok = torch.from_numpy(g["tokens"]).double().requires_grad_(True)
    params = make_tag_params(50, 512, gain=4.0, dtype=torch.float64)
    logits = oracle.tag_head_forward(tok, params)
    closs = oracle.clip_loss_single(I, T, s)
    tloss = oracle.asymmetric_loss(logits, torch.from_numpy(g["additional"]).double().repeat(1, 2))
    (closs + tloss).backward()
    tol = dict(rtol=2e-4, atol=2e-5)                             # the reference ran in fp32
    np.testing.assert_allclose(logits.detach().numpy(), g["tag_logits"], **tol)
    np.testing.assert_allclose(closs.item(), g["contrastive_loss"], rtol=1e-5)
    np.testing.assert_allclose(tloss.item(), g["tag_loss"], rtol=1e-4)
    np.testing.assert_allclose(I.grad.numpy(), g["d_image_features"], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(T.grad.numpy(), g["d_text_features"], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(tok.grad[:2, :4, :16].numpy(), g["d_tokens_head"], rtol=2e-3, atol=1e-5)
    idx = oracle.control_word_indices(logits.detach())
    tl = list(g["tag_list"])
    assert [",".join(tl[i] for i in row) for row in idx.tolist()] == list(g["words"])


def test_siglip_oracle_matches_reference(golden_dir):
    """oracle/siglip_oracle.py against the reference's SigLipLoss (loss.py:314-448): single process and 2 / 3 ranks
    under gloo with the 'gather' exchange (tests/golden/siglip.npz, oracle/make_golden.py:golden_siglip)."""
    g = np.load(os.path.join(golden_dir, "siglip.npz"))
    for n in range(int(g["n_w1"])):
        pre = f"w1_{n}_"
        I, T = torch.from_numpy(g[pre + "I"]), torch.from_numpy(g[pre + "T"])
        lo, dI, dT, ds, db = oracle.siglip_loss_world([I], [T], float(g[pre + "scale"]), float(g[pre + "bias"]))
        assert abs(float(lo[0]) - float(g[pre + "loss"])) < 1e-10 * max(1.0, abs(float(g[pre + "loss"])))
        for a, name in ((dI[0], "dI"), (dT[0], "dT"), (ds[0], "dscale"), (db[0], "dbias")):
            assert np.abs(a.numpy() - g[pre + name]).max() < 1e-10 * max(1.0, np.abs(g[pre + name]).max())
    for W in (2, 3):
        pre = f"w{W}_"
        b = int(g[pre + "b"])
        I_all, T_all = torch.from_numpy(g[pre + "I"]), torch.from_numpy(g[pre + "T"])
        Il = [I_all[r * b:(r + 1) * b] for r in range(W)]
        Tl = [T_all[r * b:(r + 1) * b] for r in range(W)]
        lo, dI, dT, ds, db = oracle.siglip_loss_world(Il, Tl, float(g[pre + "scale"]), float(g[pre + "bias"]))
        for r in range(W):
            k = f"{pre}r{r}_"
            assert abs(float(lo[r]) - float(g[k + "loss"])) < 1e-10 * max(1.0, abs(float(g[k + "loss"])))
            for a, name in ((dI[r], "dI"), (dT[r], "dT"), (ds[r], "dscale"), (db[r], "dbias")):
                assert np.abs(a.numpy() - g[k + name]).max() < 1e-10 * max(1.0, np.abs(g[k + name]).max()), (W, r, name)
