#!/bin/bash
# onesweep tile shapes (library built with EXTRA=-DHKCSA_OS_SHAPES): ms per pass on 100 M random (u64, u32) pairs
for s in ${SHAPES:-0 1 2 3 4 5 6 7}; do
  echo "shape $s: $(HKCSA_OS_SHAPE=$s CASES=random python tools/sort_probe.py 2>&1 | tail -1)"
done
