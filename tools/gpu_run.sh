#!/bin/bash
# One parametrised script for every gpurun call (replaces the per-call scripts of round 1).
#   usage (repo root, on the GPU box):  bash tools/gpu_run.sh <tag> <step> [<step> ...]
# steps:  smoke | tests[:<pytest -k expr>] | bench[:<steps>] | ref | secondary | extra | launches | ncu[:<kernel regex>] |
#         ncuwallish | sanitizer | sass | launchpy:<script> | sh:<script> | py:<script> (python tools/lab/<script>.py, output in gpurun_out/<script>_<tag>.log)
# Everything lands in gpurun_out/ with the tag in its name; copy what should be judged into profiles/.
TAG=${1:-r00}; shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi_$TAG.csv 2>&1
for STEP in "$@"; do
  NAME=${STEP%%:*}; ARG=""; [[ "$STEP" == *:* ]] && ARG=${STEP#*:}
  echo "=== $STEP"
  case $NAME in
    smoke)
      python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke_$TAG.log; tail -2 $OUT/smoke_$TAG.log ;;
    tests)
      if [ -n "$ARG" ]; then timeout 1500 python -m pytest tests -m gpu -x -q -k "$ARG" > $OUT/pytest_$TAG.log 2>&1
      else timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; fi
      echo "pytest rc=$?" >> $OUT/pytest_$TAG.log; tail -15 $OUT/pytest_$TAG.log ;;
    bench)
      timeout 900 python bench.py ${ARG:+--steps $ARG} > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err ;;
    bench20)
      timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench20_$TAG.json 2> $OUT/bench20_$TAG.err; echo "bench20 rc=$?"; cut -c1-900 $OUT/bench20_$TAG.json ;;
    ref)
      timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"; cat $OUT/bench_ref_$TAG.json; tail -3 $OUT/bench_ref_$TAG.err ;;
    secondary)
      timeout 1200 python bench.py --secondary > $OUT/secondary_$TAG.json 2> $OUT/secondary_$TAG.err; echo "secondary rc=$?"; cat $OUT/secondary_$TAG.json; tail -3 $OUT/secondary_$TAG.err ;;
    extra)
      timeout 900 python tools/bench_extra.py > $OUT/extra_$TAG.json 2> $OUT/extra_$TAG.err; cat $OUT/extra_$TAG.json ;;
    launches)
      CPF_BENCH_QUICK=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_$TAG.csv \
          python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1; tail -3 $OUT/launches_$TAG.csv ;;
    ncu)
      CPF_BENCH_QUICK=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:${ARG:-fftlog_stream} -s 4 -c 1 -f -o $OUT/prof_$TAG \
          python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1; tail -2 $OUT/ncu_full_$TAG.log ;;
    ncuwallish)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:${ARG:-wallish} -s 2 -c 3 -f -o $OUT/prof_wallish_$TAG \
          python tools/lab/wallish_run.py 8192 > $OUT/ncu_wallish_$TAG.log 2>&1; tail -2 $OUT/ncu_wallish_$TAG.log
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_wallish_$TAG.csv \
          python tools/lab/wallish_run.py 8192 >> $OUT/ncu_wallish_$TAG.log 2>&1 ;;
    sanitizer)
      timeout 1500 compute-sanitizer --tool ${ARG:-memcheck} python -m pytest tests -m gpu -x -q -k "golden or persistent or wallish" > $OUT/sanitizer_${ARG:-memcheck}_$TAG.log 2>&1
      tail -5 $OUT/sanitizer_${ARG:-memcheck}_$TAG.log ;;
    sass)
      bash tools/sass_summary.sh > $OUT/sass_$TAG.txt 2>&1; cat $OUT/sass_$TAG.txt ;;
    ncupy)   # ncupy:<script>,<kernel regex>[,<skip>]: ncu --set full of one launch of a kernel in tools/lab/<script>.py
      IFS=, read SCRIPT REGEX SKIP <<< "$ARG"
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:$REGEX -s ${SKIP:-3} -c 1 -f -o $OUT/prof_${SCRIPT}_$TAG \
          python tools/lab/$SCRIPT.py > $OUT/ncu_${SCRIPT}_$TAG.log 2>&1; tail -2 $OUT/ncu_${SCRIPT}_$TAG.log ;;
    launchpy)   # launchpy:<script>: launch list (kernel names + durations) of tools/lab/<script>.py
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_${ARG}_$TAG.csv \
          python tools/lab/$ARG.py > $OUT/ncu_launch_${ARG}_$TAG.log 2>&1; tail -2 $OUT/ncu_launch_${ARG}_$TAG.log
      python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('$OUT/launches_${ARG}_$TAG.csv')) if len(r) > 10 and r[0].isdigit()]
c = collections.Counter(); t = collections.Counter()
for r in rows:
    name = r[4].split('(')[0][:90]; c[name] += 1; t[name] += float(r[-1].replace(',', ''))
for n, _ in t.most_common(40): print('%6d x %10.1f us  %s' % (c[n], t[n] / 1e3 if False else t[n], n))
PY
      ;;
    sh)      # sh:<script>: bash tools/lab/<script>.sh, output in gpurun_out/<script>_<tag>.log
      timeout 1200 bash tools/lab/$ARG.sh > $OUT/${ARG}_$TAG.log 2>&1; echo "rc=$?" >> $OUT/${ARG}_$TAG.log; tail -40 $OUT/${ARG}_$TAG.log ;;
    py)
      timeout 1200 python tools/lab/$ARG.py > $OUT/${ARG}_$TAG.log 2>&1; echo "rc=$?" >> $OUT/${ARG}_$TAG.log; tail -40 $OUT/${ARG}_$TAG.log ;;
    *) echo "unknown step $STEP" ;;
  esac
done
ls $OUT | grep $TAG
