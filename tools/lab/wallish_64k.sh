#!/bin/bash
# Wallish2018 at the BASELINE size (65 536 spectra) with the phase split (CPF_WALLISH_DBG) and the plain timing
CPF_WALLISH_DBG=1 python tools/lab/wallish_run.py 65536 2 2>&1 | grep -m1 resident
CPF_WALLISH_DBG=1 python tools/lab/wallish_run.py 65536 2 2>&1 | tail -3
python tools/lab/wallish_run.py 65536 5 2>&1 | tail -1
