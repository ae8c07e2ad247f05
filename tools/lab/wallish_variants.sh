# lab: timing of the compile-time variants of wallish_fused_kernel (CPF_WALLISH_VARIANT bits: 1 shared FFT code, 2 rolled log, 4 rolled exp)
for p in 0 1; do for v in 0 1 2 3 4 5 6 7; do echo "== persistent=$p variant=$v"; CPF_WALLISH_PERSISTENT=$p CPF_WALLISH_VARIANT=$v python tools/lab/wallish_run.py 65536 5; done; done
