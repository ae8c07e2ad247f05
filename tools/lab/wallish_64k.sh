#!/bin/bash
# Wallish2018 at the BASELINE size (65 536 spectra): column layout (cpf_wallish2018) against one row per spectrum (cpf_wallish2018_rows)
python tools/lab/wallish_run.py 65536 5 2>&1 | tail -1
python tools/lab/wallish_run.py 65536 5 rows 2>&1 | tail -2
python tools/lab/wallish_run.py 1023 3 rows 2>&1 | tail -2
