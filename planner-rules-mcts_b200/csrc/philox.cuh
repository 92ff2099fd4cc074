// Counter-based Philox4x32-10 (Salmon et al., SC'11) for reproducible rollouts.
// The reference draws rollout actions from a process-global std::mt19937
// (/root/reference/src/behavior_rules_mcts.hpp:154); a counter-based generator gives
// every (rollout, step, agent) its own draw without any shared state.
#pragma once
#include <stdint.h>

namespace brm {

struct Philox4 {
  uint32_t w[4];
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 out;
  out.w[0] = c0;
  out.w[1] = c1;
  out.w[2] = c2;
  out.w[3] = c3;
  return out;
}

// One block of four draws serves steps 4b..4b+3 of (rollout, agent).
__device__ __forceinline__ Philox4 action_block(uint64_t seed, uint64_t rollout, uint32_t step,
                                                uint32_t agent) {
  return philox4x32_10(static_cast<uint32_t>(rollout), static_cast<uint32_t>(rollout >> 32),
                       step >> 2, agent, static_cast<uint32_t>(seed),
                       static_cast<uint32_t>(seed >> 32));
}

__device__ __forceinline__ uint32_t pick_word(const Philox4& b, uint32_t step) {
  const uint32_t s = step & 3u;
  return s == 0 ? b.w[0] : (s == 1 ? b.w[1] : (s == 2 ? b.w[2] : b.w[3]));
}

__device__ __forceinline__ uint32_t draw_action(uint32_t word, uint32_t n_actions) {
  return __umulhi(word, n_actions);
}

}  // namespace brm
