for cfg in "0 0" "150 1" "250 1" "400 1" "250 0" "0 0"; do
  set -- $cfg
  echo -n "first-barrier offset $1 ns inv $2: "
  CPF_STREAM_WSKEW_NS=$1 CPF_STREAM_WSKEW_INV=$2 python tools/lab/stream_probe.py 2>&1 | tail -1
done
