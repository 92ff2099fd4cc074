#!/bin/bash
# SASS evidence per kernel of libbrm.so: counts of TMA (UBLKCP), mbarrier (SYNCS), named-barrier
# (BAR) and FP64 opcodes.  usage: tools/sass_summary.sh > profiles/rNN_sass_summary.txt
SO=${1:-planner-rules-mcts_b200/libbrm.so}
echo "# SASS evidence of the TMA staging path (cuobjdump -sass $SO)"
echo "# UBLKCP = cp.async.bulk (non-tensor TMA) global -> shared; SYNCS = mbarrier; no UTMALDG / UTCMMA: nothing here is a contraction"
cuobjdump -sass "$SO" | awk '
  /Function :/ { f=$3; next }
  { for (i = 1; i <= NF; i++) if ($i ~ /^(UBLKCP|SYNCS|UTMALDG|UTCMMA|DADD|DMUL|DFMA|BAR)/) { split($i, a, "."); c[f " " a[1]]++ } }
  END { for (k in c) print c[k], k }' | sort -k2 | while read n f op; do echo "$n $(echo $f | c++filt) $op"; done
