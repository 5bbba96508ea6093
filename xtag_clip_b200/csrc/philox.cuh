// Philox4x32-10 counter-based RNG for the attention-probability dropout of the tag head
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

__host__ __device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// true = keep.  One Philox block serves 4 consecutive element indices.
__host__ __device__ __forceinline__ bool philox_keep(uint64_t seed, uint64_t offset, uint64_t idx, float p_drop) {
  uint32_t r[4];
  philox4x32_10(seed, idx >> 2, offset, r);
  const uint32_t u = r[idx & 3];
  return (float)(u >> 8) * (1.0f / 16777216.0f) >= p_drop;
}

}  // namespace xtag
