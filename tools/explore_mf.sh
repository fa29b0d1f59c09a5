#!/bin/bash
# config 3 under different separator grouping / leaf sizes (DIFFOPT_B200_MF_GROUP, DIFFOPT_B200_MF_LEAF)
for g in 1 2 3; do for l in 24 32; do
  echo -n "group $g leaf $l: "
  DIFFOPT_B200_MF_GROUP=$g DIFFOPT_B200_MF_LEAF=$l python bench_aux.py --configs 3nocpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('fronts', d['fronts'], 'levels', d['tree_levels'], 'max front', d['largest_front'], 'factor ms %.3f' % d['factor_device_ms'], 'solve ms %.3f' % d['solve_device_ms'], 'resid %.1e' % d['max_rel_residual_first_8_columns'])"
done; done
