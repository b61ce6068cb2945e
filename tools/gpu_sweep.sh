#!/bin/bash
# quick parameter sweeps on the GPU box: prints the stage times of the 3rd proof for each setting
for v in 0 1 2; do echo "quotient variant $v"; ZKIR_QUOTIENT_VARIANT=$v python tools/prove_once.py 3 | tail -2 | head -1 | grep -oE "'quotient': [0-9.]+"; done
for b in 4 8 16 28 56 112; do echo "lde batch $b"; ZKIR_LDE_BATCH=$b python tools/prove_once.py 3 | tail -2 | head -1 | grep -oE "'lde': [0-9.]+"; done
