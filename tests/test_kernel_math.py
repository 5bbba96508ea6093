"""Numerical model (numpy, fp32 arithmetic with fp32 exp2/log2) of the two epilogue algorithms of the tcgen05 kernels,
checked against exact fp64 math on adversarial inputs.  This is the CPU-side statement of WHY the fast paths and their
range tests are safe (csrc/clip_tc.cu: lse_tile, ds_tile); the kernels themselves are tested on the GPU.

  K1  per 32x32 chunk: ONE exponential per element, referenced to the row's chunk maximum cm_i; the column sums reuse
      those exponentials scaled by 2^(cm_i - W), W = max_i cm_i -- exact while the chunk spans <= kRange = 120 log2
      units (else the kernel takes the two-exponential path).
  K2  dS_ij = 2^(v_ij - rl2_i) * (g w_r + (g w_c 2^(rl2_i - nu)) 2^(nu - cl2_j)); exact while |rl2_i - cl2_j| <= 60
      over a warp's 32x128 block (else: two exponentials with the weights folded into the offsets).
"""
import numpy as np
import pytest

f32 = np.float32
K1_RANGE = f32(120.0)
K2_RANGE = f32(60.0)


def exp2(x):
    with np.errstate(over="ignore", under="ignore"):
        y = np.exp2(x.astype(np.float64)).astype(f32)
    y[np.abs(y) < np.finfo(f32).tiny] = 0.0          # ex2.approx.ftz flushes denormals
    return y


def k1_chunk_fast(v):
    """v [32, 32] fp32 log2-domain logits of one chunk -> (row (max, sum), column lse2) the fast path produces."""
    cm = v.max(axis=1)                                # per-row chunk maximum
    e = exp2(v - cm[:, None])                         # the one exponential per element
    rsum = e.sum(axis=1, dtype=f32)
    W = cm.max()
    f = exp2(cm - W)
    colsum = (e * f[:, None]).sum(axis=0, dtype=f32)
    with np.errstate(divide="ignore"):
        col = np.where(colsum > 0, W + np.log2(colsum.astype(np.float64)).astype(f32), -np.inf).astype(f32)
    return cm, rsum, col


def k1_fast_allowed(v):
    return (v.max(axis=1).max() - v.min()) <= K1_RANGE


@pytest.mark.parametrize("spread", [0.0, 5.0, 40.0, 100.0, 119.0])
def test_k1_single_exponential_column_lse_is_exact_within_its_range(spread):
    rng = np.random.default_rng(int(spread))
    worst = 0.0
    for _ in range(50):
        v = (rng.standard_normal((32, 32)) * 3 + rng.uniform(-spread / 2, spread / 2, (32, 1))).astype(f32)
        v = np.clip(v, v.max() - spread - 0.0, None).astype(f32) if spread else v
        if not k1_fast_allowed(v):
            continue
        cm, rsum, col = k1_chunk_fast(v)
        ref_col = np.log2(np.exp2(v.astype(np.float64) - v.max()).sum(axis=0)) + v.max()
        ref_row = np.log2(np.exp2(v.astype(np.float64) - v.max(axis=1, keepdims=True)).sum(axis=1)) + v.max(axis=1)
        row = cm + np.log2(rsum.astype(np.float64))
        worst = max(worst, np.abs(col - ref_col).max(), np.abs(row - ref_row).max())
    assert worst < 2e-5                               # log2 units; the loss tolerance is 1e-3 relative


def test_k1_range_test_rejects_chunks_where_the_fast_path_would_lose_a_column():
    """A column dominated by an element 2^-130 below the block reference: the shared exponentials flush it to zero.
    The range test must send such a chunk to the exact path."""
    v = np.full((32, 32), -200.0, dtype=f32)
    v[:, 0] = 0.0                                     # every row's chunk maximum is 0 -> W = 0
    v[5, 7] = -130.0                                  # column 7's LSE is ~ -130, far below W
    assert not k1_fast_allowed(v)
    _, _, col = k1_chunk_fast(v)
    ref = np.log2(np.exp2(v.astype(np.float64)).sum(axis=0))
    assert abs(col[7] - ref[7]) > 1.0 or not np.isfinite(col[7])      # the fast path really is wrong here
    # within the range it is right
    v2 = v.copy()
    v2[v2 == -200.0] = -110.0
    v2[5, 7] = -100.0
    assert k1_fast_allowed(v2)
    _, _, col2 = k1_chunk_fast(v2)
    ref2 = np.log2(np.exp2(v2.astype(np.float64)).sum(axis=0))
    assert np.abs(col2 - ref2).max() < 1e-4


def k2_block(v, rl2, cl2, g, w_r, w_c, fast):
    """dS of one warp block (rows x 128 columns) as the kernel computes it."""
    if fast:
        nu = cl2[0]
        C = exp2(nu - cl2)
        R = (f32(g) * f32(w_c) * exp2(rl2 - nu)).astype(f32)
        e = exp2(v - rl2[:, None])
        return (e * (R[:, None] * C[None, :] + f32(g) * f32(w_r)).astype(f32)).astype(f32)
    lr, lc = f32(np.log2(w_r)) if w_r > 0 else f32(-np.inf), f32(np.log2(w_c)) if w_c > 0 else f32(-np.inf)
    return (f32(g) * (exp2(v - (rl2 - lr)[:, None]) + exp2(v - (cl2 - lc)[None, :]))).astype(f32)


def k2_fast_allowed(rl2, cl2):
    return (rl2.max() - cl2.min() <= K2_RANGE) and (cl2.max() - rl2.min() <= K2_RANGE)


@pytest.mark.parametrize("row_shift", [0.0, 0.3, 0.6, 1.0])
@pytest.mark.parametrize("scale", [14.285714, 100.0])
def test_k2_one_exponential_matches_exact_ds_inside_the_range(row_shift, scale):
    rng = np.random.default_rng(int(row_shift * 10 + scale))
    M = N = 128
    a = (rng.standard_normal((M, N)) * 0.05 + rng.uniform(-row_shift, row_shift, (M, 1))).astype(f32)
    a[np.arange(M), np.arange(M)] += 0.5              # matched pairs
    sl2 = f32(scale * 1.4426950408889634)
    v = (a * sl2).astype(f32)
    v64 = v.astype(np.float64)
    rl2 = (np.log2(np.exp2(v64 - v64.max(1, keepdims=True)).sum(1)) + v64.max(1)).astype(f32)
    cl2 = (np.log2(np.exp2(v64 - v64.max(0, keepdims=True)).sum(0)) + v64.max(0)).astype(f32)
    g, w = 0.7, 0.5 / M
    exact = g * w * (np.exp2(v64 - rl2[:, None].astype(np.float64)) + np.exp2(v64 - cl2[None, :].astype(np.float64)))
    for r0 in range(0, M, 32):                        # one warp = 32 rows x 128 columns
        blk = slice(r0, r0 + 32)
        fast = k2_fast_allowed(rl2[blk], cl2)
        d = k2_block(v[blk], rl2[blk], cl2, g, w, w, fast)
        assert np.isfinite(d).all()
        # bf16 storage rounds each element to 2^-9 relative; the formula itself must be far inside that
        err = np.abs(d - exact[blk]).max() / exact[blk].max()
        assert err < 2e-5, (fast, err)


def test_k2_fast_path_never_overflows_at_the_edge_of_its_range():
    """Worst case the range test still admits: rl2 - cl2 = +-60 in one block."""
    rl2 = np.array([0.0, 60.0] * 16, dtype=f32)       # rows 60 apart
    cl2 = np.concatenate([np.zeros(64), np.full(64, 60.0)]).astype(f32)
    assert k2_fast_allowed(rl2, cl2)
    v = np.minimum(rl2[:, None], cl2[None, :]).astype(f32) - f32(1.0)      # S <= both LSEs, as for real logits
    d = k2_block(v, rl2, cl2, 1.0, 0.5, 0.5, True)
    ref = 0.5 * (np.exp2(v.astype(np.float64) - rl2[:, None]) + np.exp2(v.astype(np.float64) - cl2[None, :]))
    assert np.isfinite(d).all() and np.abs(d - ref).max() <= 1e-6 * ref.max()
    # one step beyond: the test refuses, and the two-exponential path is exact there
    cl2b = cl2.copy()
    cl2b[0] = -1.0
    assert not k2_fast_allowed(rl2, cl2b)
    vb = np.minimum(rl2[:, None], cl2b[None, :]).astype(f32) - f32(1.0)
    d2 = k2_block(vb, rl2, cl2b, 1.0, 0.5, 0.5, False)
    ref2 = 0.5 * (np.exp2(vb.astype(np.float64) - rl2[:, None]) + np.exp2(vb.astype(np.float64) - cl2b[None, :]))
    assert np.abs(d2 - ref2).max() <= 1e-6 * ref2.max()


def test_k2_zero_column_weight_and_infinite_column_lse():
    """local_loss without gather_with_grad: w_col = 0 and col_lse = +inf must give exactly the row term, no NaN."""
    rng = np.random.default_rng(1)
    v = (rng.standard_normal((32, 128)) * 2).astype(f32)
    rl2 = (np.log2(np.exp2(v.astype(np.float64)).sum(1))).astype(f32)
    cl2 = np.full(128, np.inf, dtype=f32)
    with np.errstate(invalid="ignore"):
        e = exp2(v - rl2[:, None])
        d = (e * f32(0.25)).astype(f32)               # fast path with has_col == False: factor = g * w_row only
    d_slow = k2_block(v, rl2, cl2, 0.5, 0.5, 0.0, False)
    assert np.isfinite(d).all() and np.isfinite(d_slow).all()
    assert np.abs(d - d_slow).max() <= 1e-6 * d.max()
