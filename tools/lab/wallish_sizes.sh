#!/bin/bash
# Wallish2018 throughput and phase split against the number of spectra: 1024 / 2048 columns keep pklin (32 KB per spectrum) resident in L2
# across repetitions, 8192+ stream it from HBM -- separates the cost of the strided gather from the arithmetic.
for n in 1024 2048 4096 8192 32768; do
  CPF_WALLISH_DBG=1 python tools/lab/wallish_run.py $n 5 2>&1 | tail -2
done
