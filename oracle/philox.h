/* TEST INFRASTRUCTURE -- oracle-side Philox4x32-10 (Salmon et al., SC'11,
 * "Parallel random numbers: as easy as 1, 2, 3"; Random123 constants).  The product
 * has its own device implementation in planner-rules-mcts_b200/csrc/philox.cuh; this
 * copy exists so the CPU oracles can replay exactly the action sequences a device
 * rollout draws.  Known answers checked in tests/test_philox.py. */
#ifndef ORACLE_PHILOX_H_
#define ORACLE_PHILOX_H_

#include <stdint.h>

static inline void oracle_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2],
                                        uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; ++r) {
    if (r > 0) {
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* action draw shared by all oracles: word (step & 3) of block (rollout, step >> 2, agent) */
static inline uint32_t oracle_action_draw(uint64_t seed, uint64_t rollout, uint32_t step,
                                          uint32_t agent, uint32_t n_actions) {
  uint32_t ctr[4] = {(uint32_t)rollout, (uint32_t)(rollout >> 32), step >> 2, agent};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t out[4];
  oracle_philox4x32_10(ctr, key, out);
  return (uint32_t)(((uint64_t)out[step & 3u] * n_actions) >> 32);
}

#endif /* ORACLE_PHILOX_H_ */
