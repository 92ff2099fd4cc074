// TEST INFRASTRUCTURE (oracle shim) -- not part of the product path.
// Stand-in for `ltl/label.h` of bark-simulator/rule-monitoring @ 187c125a
// (/root/reference/util/deps.bzl:26-31): a label is (name, optional agent id),
// ordered so it can key a std::map (gridworld/grid_world_state.cpp:47,71,83).
#ifndef ORACLE_SHIM_LTL_LABEL_H_
#define ORACLE_SHIM_LTL_LABEL_H_

#include <map>
#include <string>
#include <tuple>

namespace ltl {

class Label {
 public:
  Label() : agent_id_(-1), agent_specific_(false) {}
  Label(const std::string& s)  // NOLINT: implicit, the reference writes Label("x")
      : str_(s), agent_id_(-1), agent_specific_(false) {}
  Label(const char* s) : str_(s), agent_id_(-1), agent_specific_(false) {}  // NOLINT
  Label(const std::string& s, int agent_id)
      : str_(s), agent_id_(agent_id), agent_specific_(true) {}
  const std::string& GetLabelStr() const { return str_; }
  int GetAgentId() const { return agent_id_; }
  bool IsAgentSpecific() const { return agent_specific_; }
  bool operator<(const Label& o) const {
    return std::tie(str_, agent_specific_, agent_id_) <
           std::tie(o.str_, o.agent_specific_, o.agent_id_);
  }
  bool operator==(const Label& o) const {
    return str_ == o.str_ && agent_specific_ == o.agent_specific_ &&
           agent_id_ == o.agent_id_;
  }

 private:
  std::string str_;
  int agent_id_;
  bool agent_specific_;
};

typedef std::map<Label, bool> EvaluationMap;

}  // namespace ltl

#endif  // ORACLE_SHIM_LTL_LABEL_H_
