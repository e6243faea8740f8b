#!/bin/bash
# host entry point: chunk sizes between 2^19 and 2^20 on one box, alternating
for rep in 1 2 3; do for C in 524288 655360 786432 1048576; do
  python tools/diag_e2e.py --only-chunk $C 2>&1 | grep "chunk $C" | sed "s/side streams default plan-stream priority True slots 4 //"
done; done
