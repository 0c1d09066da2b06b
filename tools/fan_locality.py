#!/usr/bin/env python
"""How local is the element order of an NVB mesh?  
For S(level): the share of nodes whose whole fan lies inside one tile of C consecutive
elements, and the distribution of the fan span (last - first incident element).

    python tools/fan_locality.py [level=9]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from p1afempy_b200 import meshgen

level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
c, e, _ = meshgen.unit_square(level)
m, n = e.shape[0], c.shape[0]
flat, el = e.reshape(-1), np.repeat(np.arange(m), 3)
first, last = np.full(n, m), np.full(n, -1)
np.minimum.at(first, flat, el)
np.maximum.at(last, flat, el)
print('S(%d): M=%d N=%d' % (level, m, n))
for tile in (256, 1024, 4096, 16384, 65536):
    same = (first // tile) == (last // tile)
    print('tile %6d: %.1f %% of the nodes complete, %.1f %% of the incidences'
          % (tile, 100 * same.mean(), 100 * same[flat].mean()))
print('fan span percentiles 50/90/99/100:', np.percentile(last - first, [50, 90, 99, 100]))
ids = np.unique(e[:4096].reshape(-1))
print('tile 0 of 4096 elements: %d distinct nodes in %d runs of consecutive ids'
      % (ids.size, int(np.sum(np.diff(ids) != 1)) + 1))
