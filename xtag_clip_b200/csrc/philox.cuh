// Philox4x32 counter-based RNG for the attention-probability dropout of the tag head
// (reference: nn.Dropout(attention_probs_dropout_prob=0.1), tagging_heads/bert.py:168, 255).
// The keep-mask is a pure function of (seed, offset, element index), so forward and backward
// regenerate the same mask without storing it.  Statistically equivalent to torch's dropout,
// not bit-identical (torch's stream layout is an implementation detail of ATen).
#pragma once
#include <stdint.h>

namespace xtag {

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// kRounds = 7: the shortest Philox4x32 variant that passes BigCrush (Salmon et al., SC'11); a dropout mask needs no more
constexpr int kPhiloxRounds = 7;
__host__ __device__ __forceinline__ void philox4x32(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < kPhiloxRounds; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// Mask layout (shared by every K4 kernel, forward and backward): the attention probability of (row_id, key), with
// row_id = (sample*heads + head)*Lq + query, is kept iff component (key & 3) of the Philox block with counter
// row_id * kblocks + (key >> 2), kblocks = ceil(Lk / 4), has its top 24 bits >= thr = ceil(p_drop * 2^24).
// One block therefore serves 4 consecutive keys of one query row.
__host__ __device__ __forceinline__ uint32_t philox_drop_threshold(float p_drop) {
  const float t = p_drop * 16777216.0f;
  uint32_t thr = (uint32_t)t;
  if ((float)thr < t) ++thr;
  return thr;
}
// keep-bits of the 4 keys of block (row_id, kblk): bit c = key 4*kblk + c is kept
__host__ __device__ __forceinline__ uint32_t philox_keep4(uint64_t seed, uint64_t offset, uint64_t row_id, int kblk,
                                                          int kblocks, uint32_t thr) {
  uint32_t r[4];
  philox4x32(seed, row_id * (uint64_t)kblocks + (uint64_t)kblk, offset, r);
  return ((r[0] >> 8) >= thr ? 1u : 0u) | ((r[1] >> 8) >= thr ? 2u : 0u) | ((r[2] >> 8) >= thr ? 4u : 0u) |
         ((r[3] >> 8) >= thr ? 8u : 0u);
}
// true = keep (scalar form: one block per call)
__host__ __device__ __forceinline__ bool philox_keep(uint64_t seed, uint64_t offset, uint64_t row_id, int key,
                                                     int kblocks, uint32_t thr) {
  return (philox_keep4(seed, offset, row_id, key >> 2, kblocks, thr) >> (key & 3)) & 1u;
}

}  // namespace xtag
