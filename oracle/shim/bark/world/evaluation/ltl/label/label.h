// TEST INFRASTRUCTURE (oracle shim) -- not part of the product path.
// Stand-in for bark's `bark/world/evaluation/ltl/label/label.h` (bark @ 9649bc11,
// /root/reference/util/deps.bzl:6-11).  The reference uses `Label(str)` and
// `Label(str, agent_id)` as map keys
// (gridworld/label_evaluator/evaluator_label_base.h:14-28).
#ifndef ORACLE_SHIM_BARK_LABEL_H_
#define ORACLE_SHIM_BARK_LABEL_H_

#include "ltl/label.h"

namespace bark {
namespace world {
namespace evaluation {
using ltl::Label;
}  // namespace evaluation
}  // namespace world
}  // namespace bark

#endif  // ORACLE_SHIM_BARK_LABEL_H_
