#!/bin/bash
mkdir -p gpurun_out
for box in 1 4; do
  echo "== G4BOX=$box"
  B2PC_CONV_G4BOX=$box timeout 90 python tools/conv_ta_check.py 2>&1 | tail -14 | cut -c1-160
done
