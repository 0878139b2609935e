#!/bin/bash
# LayerNorm backward A/B (row-owner kernel vs two-phase kernel), alternating so both see the same clock state
for i in 1 2 3; do
  echo "rows:";   python tools/ln_bench.py | grep bwd
  echo "phases:"; CDR_LN_BWD=phases python tools/ln_bench.py | grep bwd
done
