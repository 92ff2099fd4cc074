// TEST INFRASTRUCTURE (oracle shim) -- not part of the product path.
//
// Stand-in for `ltl/rule_monitor.h` + `ltl/rule_state.h` of
// bark-simulator/rule-monitoring @ 187c125a18979214d638ca771dd86e7934932b94
// (/root/reference/util/deps.bzl:26-31).  The real library turns an LTL formula
// into a deterministic automaton with Spot; neither is available here, so this
// shim is a TABLE-DRIVEN deterministic monitor with the call contract visible at
// the reference's call sites:
//   float Evaluate(const EvaluationMap&, RuleState&) const   gridworld/grid_world_state.cpp:77-78
//   float FinalTransit(const RuleState&) const                gridworld/grid_world_state.cpp:147-148
//   RuleState::GetPriority/GetAutomaton/GetAgentIds           gridworld/grid_world_state.cpp:76-78
//   MakeRuleState(new_ids, existing_ids)                      gridworld/base_test_env.cpp:133
// Semantics ("parity unpinned" beyond test/single_agent_test.cpp:143,159):
//   * letter = sum_k value(ap_k) << (n_aps-1-k); agent-specific APs (`name#0`) are
//     looked up as Label(name, state.agent_ids[0]).
//   * table entry 0xFF = no transition (bad prefix): return weight, ++violated,
//     go to the initial state (policy reset) or stay (policy stay).
//   * FinalTransit: 0 when the current state accepts the empty suffix, else weight.
#ifndef ORACLE_SHIM_LTL_RULE_MONITOR_H_
#define ORACLE_SHIM_LTL_RULE_MONITOR_H_

#include <cstdint>
#include <map>
#include <memory>
#include <ostream>
#include <string>
#include <utility>
#include <vector>

#include "ltl/label.h"

namespace ltl {

class RuleMonitor;

class RuleState {
 public:
  RuleState(int init_state, std::shared_ptr<const RuleMonitor> automaton,
            std::vector<int> agent_ids, int priority)
      : current_state_(init_state),
        violated_(0),
        automaton_(std::move(automaton)),
        agent_ids_(std::move(agent_ids)),
        priority_(priority) {}
  int GetPriority() const { return priority_; }
  const std::shared_ptr<const RuleMonitor>& GetAutomaton() const { return automaton_; }
  const std::vector<int>& GetAgentIds() const { return agent_ids_; }
  int GetCurrentState() const { return current_state_; }
  unsigned GetViolationCount() const { return violated_; }

  int current_state_;
  unsigned violated_;

 private:
  std::shared_ptr<const RuleMonitor> automaton_;
  std::vector<int> agent_ids_;
  int priority_;
};

inline std::ostream& operator<<(std::ostream& os, const RuleState& rs) {
  os << "RuleState(q=" << rs.current_state_ << ", violated=" << rs.violated_ << ")";
  return os;
}

struct TableAp {
  std::string name;
  bool agent_specific;
};

class RuleMonitor : public std::enable_shared_from_this<RuleMonitor> {
 public:
  typedef std::shared_ptr<RuleMonitor> RuleMonitorSPtr;
  static const uint8_t kViolation = 0xFF;

  static RuleMonitorSPtr MakeTableRule(std::vector<TableAp> aps, int n_states,
                                       std::vector<uint8_t> table,
                                       std::vector<uint8_t> accepting, float weight,
                                       int priority, int init_state,
                                       bool reset_on_violation) {
    RuleMonitorSPtr r(new RuleMonitor());
    r->aps_ = std::move(aps);
    r->n_states_ = n_states;
    r->table_ = std::move(table);
    r->accepting_ = std::move(accepting);
    r->weight_ = weight;
    r->priority_ = priority;
    r->init_state_ = init_state;
    r->reset_on_violation_ = reset_on_violation;
    return r;
  }

  bool IsAgentSpecific() const {
    for (const auto& ap : aps_)
      if (ap.agent_specific) return true;
    return false;
  }

  // One rule state per "other" agent for rules with a `#0` placeholder, one
  // in total otherwise (k-permutations with k = 1 is all the reference needs:
  // gridworld/base_test_env.cpp:16-19 only uses `#0`).
  std::vector<RuleState> MakeRuleState(const std::vector<int>& new_agent_ids = {},
                                       const std::vector<int>& existing_agent_ids = {}) const {
    (void)existing_agent_ids;
    std::vector<RuleState> out;
    if (!IsAgentSpecific()) {
      out.emplace_back(init_state_, shared_from_this(), std::vector<int>(), priority_);
    } else {
      for (int id : new_agent_ids)
        out.emplace_back(init_state_, shared_from_this(), std::vector<int>{id}, priority_);
    }
    return out;
  }

  float Evaluate(const EvaluationMap& labels, RuleState& state) const {
    const int n_aps = static_cast<int>(aps_.size());
    int letter = 0;
    for (int k = 0; k < n_aps; ++k) {
      Label key = aps_[k].agent_specific ? Label(aps_[k].name, state.GetAgentIds().at(0))
                                         : Label(aps_[k].name);
      auto it = labels.find(key);
      if (it == labels.end()) return 0.0f;  // label not available: rule not evaluated
      letter |= (it->second ? 1 : 0) << (n_aps - 1 - k);
    }
    const uint8_t next = table_.at(static_cast<size_t>(state.current_state_) * (1u << n_aps) + letter);
    if (next == kViolation) {
      ++state.violated_;
      if (reset_on_violation_) state.current_state_ = init_state_;
      return weight_;
    }
    state.current_state_ = next;
    return 0.0f;
  }

  float FinalTransit(const RuleState& state) const {
    return accepting_.at(state.current_state_) ? 0.0f : weight_;
  }

 private:
  RuleMonitor() {}
  std::vector<TableAp> aps_;
  int n_states_;
  std::vector<uint8_t> table_;
  std::vector<uint8_t> accepting_;
  float weight_;
  int priority_;
  int init_state_;
  bool reset_on_violation_;
};

}  // namespace ltl

#endif  // ORACLE_SHIM_LTL_RULE_MONITOR_H_
