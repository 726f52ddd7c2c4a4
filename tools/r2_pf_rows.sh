#!/bin/bash
# same-box A/B of the L2 prefetch on the row-partitioned configurations (two CTAs per SM)
for rep in 1 2; do for pf in 0 1; do
  XH_PREFETCH=$pf timeout 200 python tools/r2_ab.py 1e9 cfg4 6; sleep 1
  XH_PREFETCH=$pf timeout 200 python tools/r2_ab.py 1e9 rows 6; sleep 1
done; done
