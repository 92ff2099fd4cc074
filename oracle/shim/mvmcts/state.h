// TEST INFRASTRUCTURE (oracle shim) -- not part of the product path.
//
// Minimal stand-in for the un-vendored dependency `mvmcts/state.h`
// (klemense1/lexmamcts @ 24ab42ca, pinned in /root/reference/util/deps.bzl:19-24).
// It only provides the names the reference's own gridworld sources use, so that
// /root/reference/gridworld/grid_world_state.cpp and gridworld/label_evaluator/*.cpp
// compile UNMODIFIED (see oracle/build_ref.sh).  The contract implemented here is
// the one visible at the reference call sites:
//   * Reward: dense double vector with Zero(n), operator()(i), rows(), operator+=
//     (used at gridworld/grid_world_state.cpp:73-81,122-130,142-149)
//   * JointAction: vector of ActionIdx (gridworld/grid_world_state.cpp:96-110)
//   * StateInterface<Impl>: CRTP base with virtual dtor and static ego_agent_idx = 0
//     (gridworld/grid_world_state.hpp:28,40; .cpp:167,175,183)
#ifndef ORACLE_SHIM_MVMCTS_STATE_H_
#define ORACLE_SHIM_MVMCTS_STATE_H_

#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <iostream>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

namespace mvmcts {

typedef std::size_t ActionIdx;
typedef std::size_t AgentIdx;
typedef std::vector<ActionIdx> JointAction;

class Reward {
 public:
  Reward() {}
  static Reward Zero(std::size_t n) {
    Reward r;
    r.v_.assign(n, 0.0);
    return r;
  }
  double& operator()(std::size_t i) { return v_.at(i); }
  const double& operator()(std::size_t i) const { return v_.at(i); }
  std::size_t rows() const { return v_.size(); }
  std::size_t size() const { return v_.size(); }
  Reward& operator+=(const Reward& o) {
    if (v_.empty()) v_.assign(o.v_.size(), 0.0);
    for (std::size_t i = 0; i < o.v_.size(); ++i) v_.at(i) += o.v_[i];
    return *this;
  }

 private:
  std::vector<double> v_;
};

typedef std::vector<Reward> JointReward;
typedef Reward ObjectiveVec;

template <class Impl>
class StateInterface {
 public:
  virtual ~StateInterface() {}
  static const AgentIdx ego_agent_idx = 0;
};

}  // namespace mvmcts

#endif  // ORACLE_SHIM_MVMCTS_STATE_H_
