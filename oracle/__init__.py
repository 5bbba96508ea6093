"""CPU oracle for the XTag-CLIP hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain PyTorch-CPU (fp64 by default) restatement of the reference's
algorithm for the path named in BASELINE.json (open_clip contrastive head + XTag
cross-attention tag head).  It is imported ONLY by ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and there only as the
checker / the timed CPU baseline.  Nothing under ``xtag_clip_b200/`` imports it; the product
path raises if the CUDA library is missing.

Parity pinning: the reference ships NO tests, fixtures or golden vectors,
so the oracle is pinned against outputs of the reference itself, executed in the build
container from /root/reference by ``oracle/make_golden.py`` (committed) and stored as small
fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks the restatement
against every fixture (fp64: <= 1e-12 abs; fp32 reference outputs: <= 2e-6 rel).
"""
from .clip_oracle import (  # noqa: F401
    l2_normalize,
    clip_logits,
    clip_loss_single,
    clip_loss_world,
    clip_loss_closed_form,
    clip_loss_local_rank,
)
from .tag_oracle import (  # noqa: F401
    TAG_CFG,
    make_tag_params,
    cross_attention_core,
    tag_head_forward,
    asymmetric_loss,
    control_word_indices,
)
from .siglip_oracle import siglip_block_loss, siglip_loss_world  # noqa: F401
from .fusion_oracle import (  # noqa: F401
    FUSION_CFG,
    make_fusion_params,
    fusion_forward,
    fusion_scores,
    dqn_cos_loss,
)
