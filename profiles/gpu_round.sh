#!/bin/bash
# One measurement round on the GPU box (run through gpurun):  profiles/gpu_round.sh TAG [tests] [bench] [launches] [full]
# Every ncu pass is preceded by the same command run plain (a number printed under ncu is never a bench value).
set -u
TAG=${1:?tag}; shift
OUT=gpurun_out
mkdir -p $OUT
for what in "$@"; do
  case $what in
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log ;;
    bench)
      timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json ;;
    kernels)
      timeout 300 python profiles/time_kernels.py > $OUT/${TAG}_time_kernels.txt 2>&1; echo "kernels rc=$?"; cat $OUT/${TAG}_time_kernels.txt ;;
    launches)
      timeout 120 python profiles/prof_step.py > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; cat $OUT/${TAG}_plain.log; continue; }
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
        python profiles/prof_step.py > $OUT/${TAG}_ncu_launches.log 2>&1; echo "launches rc=$?" ;;
    full)
      timeout 120 python profiles/prof_step.py > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; cat $OUT/${TAG}_plain.log; continue; }
      PROF_STEPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:transform_kernel -s 4 -c 4 \
        -o $OUT/${TAG}_full -f python profiles/prof_step.py > $OUT/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
      ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
      ncu -i $OUT/${TAG}_full.ncu-rep --page source --csv > $OUT/${TAG}_full_src.csv 2>/dev/null ;;
  esac
done
