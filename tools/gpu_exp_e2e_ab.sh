#!/bin/bash
# A/B on ONE box: the round-1 host pipeline (device histogram with a host wait per chunk, 2 slots, one solve stream) against the
# round-2 one (host histogram, 4 slots, a solve stream per slot), alternating, 2^20-pair chunks
for rep in 1 2 3; do
  DCOL_HOST_DEVICE_COUNT=1 DCOL_HOST_SLOTS=2 DCOL_HOST_ONE_RUN_STREAM=1 python tools/diag_e2e.py --only-chunk 1048576 2>&1 | grep "chunk 1048576" | sed "s/^/old pipeline: /"
  python tools/diag_e2e.py --only-chunk 1048576 2>&1 | grep "chunk 1048576" | sed "s/^/new pipeline: /"
done
