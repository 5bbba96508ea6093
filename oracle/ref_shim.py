"""Oracle support (test infrastructure): import the REAL reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
``oracle/make_golden.py`` to generate the fixtures under ``tests/golden/`` and by
``tests/test_oracle_vs_reference.py`` (skipped when the tree is absent).  No reference file
is modified or copied; the shims below only patch the *environment* so the unmodified
sources import under transformers 5.x:

  1. ``ftfy`` (missing) -> stub module with ``fix_text = identity`` (tokenizer.py:14)
  2. ``transformers.modeling_utils`` lost ``apply_chunking_to_forward`` /
     ``find_pruneable_heads_and_indices`` / ``prune_linear_layer`` (bert.py:39-44) -> alias them
  3. ``BertPreTrainedModel.init_weights`` / ``BertModel.get_head_mask`` need 4.x behaviour
  4. model.py:271,277 open config files relative to CWD -> chdir into src/open_clip
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("XTAG_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "open_clip", "loss.py"))


def load_ref_loss():
    """``src/open_clip/loss.py`` imports standalone (torch only)."""
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location(
        "xtag_ref_loss", os.path.join(REF_SRC, "open_clip", "loss.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref_asl():
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location(
        "xtag_ref_asl", os.path.join(REF_SRC, "open_clip", "tagging_heads", "asymmetric_loss.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _patch_transformers():
    import transformers.modeling_utils as mu
    try:
        import transformers.pytorch_utils as pu
    except Exception:  # pragma: no cover
        pu = None

    def _find_pruneable_heads_and_indices(heads, n_heads, head_size, already_pruned_heads):
        raise NotImplementedError("head pruning is not on the hot path")

    for name in ("apply_chunking_to_forward", "find_pruneable_heads_and_indices", "prune_linear_layer"):
        if not hasattr(mu, name):
            if pu is not None and hasattr(pu, name):
                setattr(mu, name, getattr(pu, name))
            elif name == "find_pruneable_heads_and_indices":
                setattr(mu, name, _find_pruneable_heads_and_indices)
            else:  # pragma: no cover
                raise ImportError(f"cannot alias transformers.{name}")


def load_ref_bert():
    """``tagging_heads/bert.py`` standalone (it has no package-relative imports)."""
    sys.dont_write_bytecode = True
    _patch_transformers()
    spec = importlib.util.spec_from_file_location(
        "xtag_ref_bert", os.path.join(REF_SRC, "open_clip", "tagging_heads", "bert.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["xtag_ref_bert"] = mod        # transformers 5.x looks the class module up
    spec.loader.exec_module(mod)
    mod.BertPreTrainedModel.init_weights = lambda self: self.apply(self._init_weights)
    mod.BertModel.get_head_mask = lambda self, hm, n, *a, **k: [None] * n
    return mod


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def load_ref_open_clip():
    """Full ``open_clip`` package of the reference (needed for CLIP.tag_forward / config 1)."""
    sys.dont_write_bytecode = True
    if "ftfy" not in sys.modules:
        stub = types.ModuleType("ftfy")
        stub.fix_text = lambda s: s
        sys.modules["ftfy"] = stub
    _patch_transformers()
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import open_clip  # noqa: F401  (the reference's, from REF_SRC)
    from open_clip.tagging_heads import bert
    bert.BertPreTrainedModel.init_weights = lambda self: self.apply(self._init_weights)
    bert.BertModel.get_head_mask = lambda self, hm, n, *a, **k: [None] * n
    return open_clip


def build_ref_tag_head(embed_dim: int):
    """The reference's tag head exactly as CLIP.__init__ builds it (model.py:270-283), without
    the encoders: returns (tag_head: BertModel, tag_labels: nn.Embedding, tag_fc: nn.Linear)."""
    import torch.nn as nn
    bert = load_ref_bert()
    cfg_path = os.path.join(REF_SRC, "open_clip", "tagging_heads", "tag_bert_config.json")
    cfg = bert.BertConfig.from_json_file(cfg_path)
    cfg.encoder_width = embed_dim
    head = bert.BertModel(config=cfg, add_pooling_layer=False)
    del head.embeddings
    for layer in head.encoder.layer:
        del layer.attention
    tag_labels = nn.Embedding(44, cfg.hidden_size)
    tag_fc = nn.Linear(cfg.hidden_size, 1)
    return head, tag_labels, tag_fc


def ref_tag_forward(head, tag_labels, tag_fc, tag_embeds):
    """Body of CLIP.tag_forward (model.py:337-352) driving the REAL BertModel."""
    import torch
    bs = len(tag_embeds)
    object_atts = torch.ones(tag_embeds.size()[:-1], dtype=torch.long)
    label_embed = tag_labels.weight.unsqueeze(0).repeat(bs, 1, 1)
    out = head(encoder_embeds=label_embed, encoder_hidden_states=tag_embeds,
               encoder_attention_mask=object_atts, return_dict=False, mode="tagging")
    return tag_fc(out[0]).squeeze(-1)
