#!/bin/bash
# stream kernel: some warps of a group held back after the first / second group barrier (CPF_STREAM_WSKEW_NS / CPF_STREAM_WSKEW2_NS / _MASK)
for cfg in "0 0 4" "0 32 4" "0 64 4" "0 100 4" "0 150 4" "0 200 4" "0 250 4" "0 300 4" "0 550 4" "100 300 4" "0 0 4"; do
  set -- $cfg
  echo -n "wskew $1 wskew2 $2 mask $3: "
  CPF_STREAM_WSKEW_NS=$1 CPF_STREAM_WSKEW2_NS=$2 CPF_STREAM_WSKEW_MASK=$3 python tools/lab/stream_probe.py 2>&1 | tail -1
done
