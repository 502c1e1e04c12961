#!/bin/bash
cd "$(dirname "$0")/.."
for w in C1 C3 C4; do
for sh in 128x4 128x16; do
  echo "== $w $sh"; NS=1000 TIKTAK_SOBOL_GROUP=0 TIKTAK_SOBOL_SHAPE=$sh python tools/sobol_fixed_cost.py $w | grep warm
done; done
