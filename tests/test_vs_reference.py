"""Runs only where the real reference tree is present (/root/reference, i.e. the build container): the oracle and the
drop-in host modules against the UNMODIFIED reference code, live (no fixtures).  Skipped on the GPU box."""
import os

import numpy as np
import pytest
import torch

import oracle
import xtag_clip_b200 as xt
from oracle import ref_shim
from oracle.tag_oracle import make_tag_params
from kernel_model import ModelKernels

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def ref_holder():
    """the reference's tag head exactly as CLIP.__init__ builds it, on a holder module with CLIP's attribute names"""
    with ref_shim._cwd(os.path.join(ref_shim.REF_SRC, "open_clip")):
        head, tag_labels, tag_fc = ref_shim.build_ref_tag_head(64)
    m = torch.nn.Module()
    m.tag_head, m.tag_labels, m.tag_fc = head, tag_labels, tag_fc
    with open(os.path.join(ref_shim.REF_SRC, "open_clip", "tagging", "scar_tag_list.txt")) as fr:
        m.tag_list = [t.strip() for t in fr.readlines()]
    m.double().eval()
    m.load_state_dict(make_tag_params(5, 64, gain=5.0, dtype=torch.float64), strict=True)
    return m


def test_reference_cliploss_live_vs_oracle_and_dropin():
    ref = ref_shim.load_ref_loss()
    g = torch.Generator().manual_seed(3)
    I = torch.nn.functional.normalize(torch.randn(40, 24, generator=g, dtype=torch.float64), dim=-1)
    T = torch.nn.functional.normalize(0.2 * I + 0.8 * torch.randn(40, 24, generator=g, dtype=torch.float64), dim=-1)
    outs = []
    for mod in (ref.ClipLoss(), xt.ClipLoss(_kernels=ModelKernels())):
        Ic, Tc = I.clone().requires_grad_(True), T.clone().requires_grad_(True)
        s = torch.tensor(25.0, dtype=torch.float64, requires_grad=True)
        loss = mod(Ic, Tc, s)
        loss.backward()
        outs.append((loss.item(), Ic.grad.numpy(), Tc.grad.numpy(), s.grad.item()))
    lo, dI, dT, ds, _, _ = oracle.clip_loss_closed_form(I, T, torch.tensor(25.0, dtype=torch.float64))
    assert rel_err(outs[0][0], lo.item()) < 1e-12 and rel_err(outs[0][1], dI.numpy()) < 1e-10
    for a, b in zip(outs[0], outs[1]):
        assert rel_err(b, a) < 1e-5          # the drop-in returns fp32 scalars / LSEs


def test_patch_reference_model_and_from_reference(ref_holder):
    """`patch_reference_model` routes the reference model's tag_forward through this package without touching its
    parameters; `TagHead.from_reference` takes the weights over under the same state_dict keys."""
    ref_shim.load_ref_open_clip()        # puts the reference's src on sys.path (with the environment shims)
    from open_clip.model import CLIP
    g = torch.Generator().manual_seed(11)
    tokens = torch.randn(3, 50, 64, generator=g, dtype=torch.float64)
    want = CLIP.tag_forward(ref_holder, tokens)
    keys_before = sorted(ref_holder.state_dict().keys())
    head = xt.TagHead.from_reference(ref_holder, _kernels=ModelKernels()).double().eval()
    assert rel_err(head.tag_forward(tokens).detach().numpy(), want.detach().numpy()) < 1e-10
    assert head.prepare_control_words(want.detach()) == CLIP.prepare_control_words(ref_holder, want.detach())
    xt.patch_reference_model(ref_holder, _kernels=ModelKernels())
    got = ref_holder.tag_forward(tokens)
    assert rel_err(got.detach().numpy(), want.detach().numpy()) < 1e-10
    assert sorted(ref_holder.state_dict().keys()) == keys_before          # no parameters / buffers added
    # gradients reach the reference model's own parameters
    got.sum().backward()
    assert ref_holder.tag_head.encoder.layer[1].crossattention.self.key.weight.grad is not None


def test_reference_asl_live(ref_holder):
    asl = ref_shim.load_ref_asl()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(5, 44, generator=g, dtype=torch.float64) * 2
    y = (torch.rand(5, 22, generator=g) > 0.6).double().repeat(1, 2)
    a = asl.AsymmetricLoss()(x, y)
    b = xt.AsymmetricLoss(_kernels=ModelKernels())(x.float(), y.float())
    assert rel_err(b.item(), a.item()) < 1e-5
