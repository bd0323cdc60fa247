#!/bin/bash
for s in s0b s1a s1b s2c s2d s3b s3c s3d hd; do
  python tools/prof_wgrad2.py $s
  for cfg in 1,2,192,1 2,1,192,1 3,1,128,1 1,2,128,1 2,2,64,1 1,2,192,2; do
    LNX_WGRAD_CFG=$cfg python tools/prof_wgrad2.py $s 2>&1 | tail -1
  done
done
