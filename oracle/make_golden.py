"""Generate tests/golden/*.npz by EXECUTING THE REAL REFERENCE (build container only).

    python -m oracle.make_golden            # writes tests/golden/

Test infrastructure.  The reference has no golden vectors of its own, so the
pins for the oracle are outputs of the unmodified reference sources under /root/reference,
run on CPU with fixed seeds.  Inputs are stored with the outputs so the fixtures are
self-contained on the GPU box (where /root/reference does not exist).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.tag_oracle import make_tag_params  # noqa: E402
from oracle.fusion_oracle import make_fusion_params  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def synth_features(seed: int, b: int, d: int, dtype=torch.float64, corr: float = 0.5):
    """Correlated, L2-normalised image/text features (SURVEY.md §8d synthetic-input recipe);
    `corr` is lowered for the high-scale cases so the loss stays far from fp32 underflow."""
    g = torch.Generator().manual_seed(seed)
    i_raw = torch.randn(b, d, generator=g, dtype=torch.float32)
    noise = torch.randn(b, d, generator=g, dtype=torch.float32)
    t_raw = corr * i_raw + (1.0 - corr) * noise
    return (F.normalize(i_raw.to(dtype), dim=-1), F.normalize(t_raw.to(dtype), dim=-1))


# ----------------------------------------------------------------------------------------------
def golden_clip_w1():
    ref = ref_shim.load_ref_loss()
    out = {}
    cases = [(16, 32, 14.285714, 0, 0.5), (48, 64, 100.0, 1, 0.12), (33, 40, 1.0, 2, 0.5), (128, 96, 30.0, 3, 0.3)]
    for n, (b, d, s, seed, corr) in enumerate(cases):
        for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            I, T = synth_features(seed, b, d, corr=corr)
            I = I.to(dt).requires_grad_(True)
            T = T.to(dt).requires_grad_(True)
            sc = torch.tensor(s, dtype=dt, requires_grad=True)
            loss = ref.ClipLoss()(I, T, sc)
            d_out = ref.ClipLoss()(I.detach(), T.detach(), sc.detach(), output_dict=True)
            assert list(d_out.keys()) == ["contrastive_loss"]
            loss.backward()
            pre = f"c{n}_{tag}_"
            out[pre + "I"] = I.detach().numpy()
            out[pre + "T"] = T.detach().numpy()
            out[pre + "scale"] = np.asarray(s, dtype=np.float64)
            out[pre + "loss"] = loss.detach().numpy()
            out[pre + "dI"] = I.grad.numpy()
            out[pre + "dT"] = T.grad.numpy()
            out[pre + "dscale"] = sc.grad.numpy()
    out["n_cases"] = np.asarray(len(cases))
    np.savez_compressed(os.path.join(OUT, "clip_w1.npz"), **out)
    print("clip_w1.npz", len(out))


def _dist_worker(rank, world, port, b, d, scale, seed, corr, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = ref_shim.load_ref_loss()
    res = {}
    for local_loss in (False, True):
        for gwg in (False, True):
            I_all, T_all = synth_features(seed, b * world, d, corr=corr)
            I = I_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
            T = T_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
            sc = torch.tensor(scale, dtype=torch.float64, requires_grad=True)
            loss = ref.ClipLoss(local_loss=local_loss, gather_with_grad=gwg, cache_labels=True,
                                rank=rank, world_size=world)(I, T, sc)
            loss.backward()
            key = f"ll{int(local_loss)}_gwg{int(gwg)}_r{rank}_"
            res[key + "loss"] = loss.detach().numpy()
            res[key + "dI"] = I.grad.numpy()
            res[key + "dT"] = T.grad.numpy()
            res[key + "dscale"] = sc.grad.numpy()
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def golden_clip_dist():
    import torch.multiprocessing as mp
    out = {}
    port = 29611
    for world, b, d, scale, seed, corr in ((2, 6, 16, 14.285714, 10, 0.5), (3, 5, 24, 60.0, 11, 0.15)):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_dist_worker, args=(r, world, port, b, d, scale, seed, corr, q))
                 for r in range(world)]
        for p in procs:
            p.start()
        got = [q.get(timeout=300) for _ in range(world)]
        for p in procs:
            p.join()
        port += 1
        pre = f"w{world}_"
        I_all, T_all = synth_features(seed, b * world, d, corr=corr)
        out[pre + "I"] = I_all.numpy()
        out[pre + "T"] = T_all.numpy()
        out[pre + "b"] = np.asarray(b)
        out[pre + "scale"] = np.asarray(scale)
        for _, res in got:
            for k, v in res.items():
                out[pre + k] = v
    np.savez_compressed(os.path.join(OUT, "clip_dist.npz"), **out)
    print("clip_dist.npz", len(out))


def _siglip_worker(rank, world, port, b, d, scale, bias, seed, impl, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = ref_shim.load_ref_loss()
    I_all, T_all = synth_features(seed, b * world, d, corr=0.4)
    I = I_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    T = T_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    sc = torch.tensor(scale, dtype=torch.float64, requires_grad=True)
    bi = torch.tensor(bias, dtype=torch.float64, requires_grad=True)
    loss = ref.SigLipLoss(rank=rank, world_size=world, dist_impl=impl)(I, T, sc, bi)
    loss.backward()
    key = f"r{rank}_"
    q.put((rank, {key + "loss": loss.detach().numpy(), key + "dI": I.grad.numpy(), key + "dT": T.grad.numpy(),
                  key + "dscale": sc.grad.numpy(), key + "dbias": bi.grad.numpy()}))
    dist.barrier()
    dist.destroy_process_group()


def golden_siglip():
    """The reference's SigLipLoss (loss.py:314-448): single process, and 2 / 3 ranks under gloo with the 'gather'
    exchange (torch.distributed.nn.all_gather; the other dist_impl variants sum the same terms)."""
    import torch.multiprocessing as mp
    ref = ref_shim.load_ref_loss()
    out = {}
    for n, (b, d, scale, bias, seed) in enumerate(((12, 16, 10.0, -10.0, 30), (33, 40, 25.0, -4.0, 31),
                                                   (8, 8, 1.0, 0.5, 32))):
        I, T = synth_features(seed, b, d, corr=0.4)
        I.requires_grad_(True)
        T.requires_grad_(True)
        sc = torch.tensor(scale, dtype=torch.float64, requires_grad=True)
        bi = torch.tensor(bias, dtype=torch.float64, requires_grad=True)
        loss = ref.SigLipLoss()(I, T, sc, bi)
        loss.backward()
        pre = f"w1_{n}_"
        out.update({pre + "I": I.detach().numpy(), pre + "T": T.detach().numpy(), pre + "scale": np.asarray(scale),
                    pre + "bias": np.asarray(bias), pre + "loss": loss.detach().numpy(), pre + "dI": I.grad.numpy(),
                    pre + "dT": T.grad.numpy(), pre + "dscale": sc.grad.numpy(), pre + "dbias": bi.grad.numpy()})
    out["n_w1"] = np.asarray(3)
    port = 29651
    for world, b, d, scale, bias, seed in ((2, 6, 16, 10.0, -10.0, 40), (3, 5, 24, 20.0, -3.0, 41)):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_siglip_worker, args=(r, world, port, b, d, scale, bias, seed, "gather", q))
                 for r in range(world)]
        for p in procs:
            p.start()
        got = [q.get(timeout=300) for _ in range(world)]
        for p in procs:
            p.join()
        port += 1
        pre = f"w{world}_"
        I_all, T_all = synth_features(seed, b * world, d, corr=0.4)
        out.update({pre + "I": I_all.numpy(), pre + "T": T_all.numpy(), pre + "b": np.asarray(b),
                    pre + "scale": np.asarray(scale), pre + "bias": np.asarray(bias)})
        for _, res in got:
            for k, v in res.items():
                out[pre + k] = v
    np.savez_compressed(os.path.join(OUT, "siglip.npz"), **out)
    print("siglip.npz", len(out))


def golden_l2norm():
    g = torch.Generator().manual_seed(20)
    x = torch.randn(9, 24, generator=g, dtype=torch.float64)
    x[2] = 0.0                       # zero row -> eps clamp (F.normalize eps=1e-12)
    x[5] = x[5] * 1e-14              # tiny row, also below eps
    x[7] = x[7] * 1e4
    x.requires_grad_(True)
    # exactly the call in model.py:313 / :333
    y = F.normalize(x, dim=-1)
    gy = torch.randn(9, 24, generator=g, dtype=torch.float64)
    y.backward(gy)
    np.savez_compressed(os.path.join(OUT, "l2norm.npz"), x=x.detach().numpy(), y=y.detach().numpy(),
                        gy=gy.numpy(), gx=x.grad.numpy())
    print("l2norm.npz")


def golden_asl():
    asl = ref_shim.load_ref_asl()
    out = {}
    g = torch.Generator().manual_seed(30)
    x = (torch.randn(7, 44, generator=g, dtype=torch.float64) * 3.0)
    x[0, 0], x[0, 1] = 40.0, -40.0   # saturate both clamps
    y22 = (torch.rand(7, 22, generator=g) > 0.7).to(torch.float64)
    y = y22.repeat(1, 2)             # train_other.py:128
    out["x"], out["y"] = x.numpy(), y.numpy()
    for n, (gn, gp, clip) in enumerate(((4, 1, 0.05), (7, 0, 0.05), (0, 0, 0.0))):
        xx = x.clone().requires_grad_(True)
        loss = asl.AsymmetricLoss(gamma_neg=gn, gamma_pos=gp, clip=clip)(xx, y)
        loss.backward()
        assert torch.is_grad_enabled()
        out[f"k{n}_cfg"] = np.asarray([gn, gp, clip], dtype=np.float64)
        out[f"k{n}_loss"] = loss.detach().numpy()
        out[f"k{n}_dx"] = xx.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "asl.npz"), **out)
    print("asl.npz")


def golden_tag_head():
    oc = ref_shim.load_ref_open_clip()
    from open_clip.model import CLIP
    out = {}
    cases = [  # (seed, D, b, N, gain)
        (40, 64, 3, 50, 1.0),
        (41, 32, 2, 7, 6.0),
        (42, 48, 2, 197, 6.0),
    ]
    for n, (seed, D, b, N, gain) in enumerate(cases):
        with ref_shim._cwd(os.path.join(ref_shim.REF_SRC, "open_clip")):
            head, tag_labels, tag_fc = ref_shim.build_ref_tag_head(D)
        params = make_tag_params(seed, D, gain=gain, dtype=torch.float64)
        holder = torch.nn.Module()
        holder.tag_head, holder.tag_labels, holder.tag_fc = head, tag_labels, tag_fc
        holder.double()
        missing, unexpected = holder.load_state_dict(params, strict=True)
        holder.eval()
        g = torch.Generator().manual_seed(seed + 1000)
        tokens = torch.randn(b, N, D, generator=g, dtype=torch.float32).double().requires_grad_(True)
        # the reference's own method body, unbound, on a holder carrying the reference modules
        logits = CLIP.tag_forward(holder, tokens)
        gl = torch.randn(b, 44, generator=g, dtype=torch.float32).double()
        logits.backward(gl)
        pre = f"t{n}_"
        out[pre + "cfg"] = np.asarray([seed, D, b, N, gain], dtype=np.float64)
        out[pre + "tokens"] = tokens.detach().numpy()
        out[pre + "logits"] = logits.detach().numpy()
        out[pre + "glogits"] = gl.numpy()
        out[pre + "dtokens"] = tokens.grad.numpy()
        out[pre + "dq0w"] = head.encoder.layer[0].crossattention.self.query.weight.grad.numpy()[:4, :8].copy()
        out[pre + "dk1w"] = head.encoder.layer[1].crossattention.self.key.weight.grad.numpy()[:4, :8].copy()
        # control words through the reference's own prepare_control_words
        with open(os.path.join(ref_shim.REF_SRC, "open_clip", "tagging", "scar_tag_list.txt")) as fr:
            tag_list = [t.strip() for t in fr.readlines()]
        holder.tag_list = tag_list
        words = CLIP.prepare_control_words(holder, logits.detach())
        out[pre + "words"] = np.asarray(words)
        out[pre + "tag_list"] = np.asarray(tag_list)
        out[pre + "state_keys"] = np.asarray(sorted(holder.state_dict().keys()))
    out["n_cases"] = np.asarray(len(cases))
    np.savez_compressed(os.path.join(OUT, "tag_head.npz"), **out)
    print("tag_head.npz")


def golden_fusion():
    """TQN fusion head + DQNCOSLoss exactly as CLIP.forward / the training loop chain them (model.py:552-561,
    train_other.py:130-132): the reference's own TQN_Model (built with a cfg object for the small widths, default
    construction for 512) with deterministic parameters, both directions, loss and gradients."""
    ref_shim.load_ref_open_clip()
    from open_clip.CAR_heads.TQN_model import TQN_Model
    asl = ref_shim.load_ref_asl()
    out = {}
    cases = [  # (seed, d_model, layers, B, image tokens, text tokens)
        (60, 64, 4, 6, 10, 7),
        (61, 32, 2, 3, 50, 5),
        (62, 512, 4, 4, 12, 9),
    ]
    for n, (seed, d, layers, B, Pi, Pt) in enumerate(cases):
        if d == 512 and layers == 4:
            model = TQN_Model()                                 # the way CLIP.__init__ builds it (model.py:286)
        else:
            cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(FUSION_DIM=d, FUSION_CLASS_NUM=1,
                                                                    FUSION_DECODER_NUM=layers))
            model = TQN_Model(cfg)
        model = model.double().eval()
        model.load_state_dict(make_fusion_params(seed, d, 1, layers), strict=True)
        g = torch.Generator().manual_seed(seed + 1000)
        out_token = torch.randn(B, Pi, d, generator=g, dtype=torch.float32).double().requires_grad_(True)
        text_tokens = torch.randn(B, Pt, d, generator=g, dtype=torch.float32).double().requires_grad_(True)
        # model.py:552-561
        text_features_l, text_features_g = text_tokens.clone(), text_tokens.clone().mean(axis=1)
        image_features_l, image_features_g = out_token.clone(), out_token.clone().mean(axis=1)
        i2t = model(torch.cat([image_features_g.unsqueeze(1), image_features_l], dim=1), text_features_g).squeeze(-1)
        t2i = model(torch.cat([text_features_g.unsqueeze(1), text_features_l], dim=1), image_features_g).squeeze(-1)
        ce = asl.DQNCOSLoss()
        l1, l2 = ce(i2t), ce(t2i)
        (l1 + l2).backward()
        pre = f"f{n}_"
        out[pre + "cfg"] = np.asarray([seed, d, layers, B, Pi, Pt], dtype=np.int64)
        out[pre + "out_token"] = out_token.detach().numpy()
        out[pre + "text_tokens"] = text_tokens.detach().numpy()
        out[pre + "i2t"] = i2t.detach().numpy()
        out[pre + "t2i"] = t2i.detach().numpy()
        out[pre + "loss_i2t"] = l1.detach().numpy()
        out[pre + "loss_t2i"] = l2.detach().numpy()
        out[pre + "d_out_token"] = out_token.grad.numpy()
        out[pre + "d_text_tokens"] = text_tokens.grad.numpy()
        out[pre + "d_inproj0"] = model.decoder.layers[0].multihead_attn.in_proj_weight.grad.numpy()[:6, :8].copy()
        out[pre + "d_mlp9"] = model.mlp_head[9].weight.grad.numpy().copy()
        out[pre + "state_keys"] = np.asarray(sorted(model.state_dict().keys()))
    # DQNCOSLoss alone on matrices with a wide dynamic range
    gx = torch.Generator().manual_seed(63)
    for n, (B, amp) in enumerate([(5, 1.0), (17, 30.0), (64, 100.0)]):
        x = (torch.randn(B, B, generator=gx, dtype=torch.float32) * amp).double().requires_grad_(True)
        l = asl.DQNCOSLoss()(x)
        l.backward()
        out[f"ce{n}_x"] = x.detach().numpy()
        out[f"ce{n}_loss"] = l.detach().numpy()
        out[f"ce{n}_dx"] = x.grad.numpy()
    out["n_cases"] = np.asarray(len(cases))
    np.savez_compressed(os.path.join(OUT, "fusion.npz"), **out)
    print("fusion.npz")


def main():
    assert ref_shim.available(), "reference tree not found"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    golden_l2norm()
    golden_asl()
    golden_clip_w1()
    golden_clip_dist()
    golden_tag_head()
    golden_fusion()
    golden_siglip()
    golden_config1()


if __name__ == "__main__" and "--siglip" in sys.argv:
    assert ref_shim.available(), "reference tree not found"
    torch.set_num_threads(4)
    golden_siglip()
elif __name__ == "__main__" and "--config1" not in sys.argv:
    main()


def golden_config1():
    """BASELINE config 1: the reference's own ViT-B-32 XTag model (random init, seed 0), batch 16 of 224x224 images
    and 77-token captions on CPU: encoders (reference code, out of scope) -> normalised features + ViT tokens, then
    the hot path exactly as the training loop runs it (train_other_simple.py:125-135): tag_forward, ClipLoss, ASL.
    The fixture stores the encoder outputs (the hot path's inputs) and the reference's head outputs / gradients."""
    oc = ref_shim.load_ref_open_clip()
    torch.manual_seed(0)
    with ref_shim._cwd(os.path.join(ref_shim.REF_SRC, "open_clip")):
        model = oc.create_model("ViT-B-32", pretrained=None, precision="fp32", device="cpu", output_dict=True)
    model.eval()
    # deterministic tag-head weights (the 54 MB of random init cannot be stored): same generator as the tests use
    params = make_tag_params(50, 512, gain=4.0, dtype=torch.float32)
    sd = model.state_dict()
    sd.update(params)
    model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(123)
    images = torch.randn(16, 3, 224, 224, generator=g)
    text = torch.randint(1, 49405, (16, 77), generator=g)
    eot = torch.randint(5, 77, (16,), generator=g)
    for i in range(16):
        text[i, eot[i]] = 49407
        text[i, eot[i] + 1:] = 0
    additional = (torch.rand(16, 22, generator=g) > 0.7).float()
    with torch.no_grad():
        image_features, tokens = model.encode_image(images, normalize=True)
        text_features, _ = model.encode_text(text, normalize=True)
    image_features = image_features.clone().requires_grad_(True)
    text_features = text_features.clone().requires_grad_(True)
    tokens = tokens.clone().requires_grad_(True)
    logit_scale = model.logit_scale.detach().exp().clone().requires_grad_(True)
    ref_loss = ref_shim.load_ref_loss()
    asl = ref_shim.load_ref_asl()
    tag_logits = model.tag_forward(tokens)
    closs = ref_loss.ClipLoss()(image_features, text_features, logit_scale)
    tloss = asl.AsymmetricLoss()(tag_logits, additional.repeat(1, 2))
    (closs + tloss).backward()
    words = model.prepare_control_words(tag_logits.detach())
    np.savez_compressed(
        os.path.join(OUT, "config1.npz"),
        image_features=image_features.detach().numpy(), text_features=text_features.detach().numpy(),
        tokens=tokens.detach().numpy().astype(np.float32), additional=additional.numpy(),
        logit_scale=logit_scale.detach().numpy(), tag_logits=tag_logits.detach().numpy(),
        contrastive_loss=closs.detach().numpy(), tag_loss=tloss.detach().numpy(),
        d_image_features=image_features.grad.numpy(), d_text_features=text_features.grad.numpy(),
        d_tokens_norm=np.asarray(tokens.grad.norm().item()), d_tokens_head=tokens.grad[:2, :4, :16].numpy().copy(),
        d_logit_scale=logit_scale.grad.numpy(), words=np.asarray(words), tag_list=np.asarray(model.tag_list),
        n_params=np.asarray(sum(p.numel() for p in model.parameters())))
    print("config1.npz", float(closs), float(tloss), words[:2])


if __name__ == "__main__" and "--config1" in sys.argv:
    os.makedirs(OUT, exist_ok=True)
    golden_config1()
