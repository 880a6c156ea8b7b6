set -u
OUT=gpurun_out; mkdir -p $OUT; T=r3j
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${T}_tests.log
for rf in 1 4 8 16 0; do
  for sz in "32 192000" "256 192000"; do set -- $sz
    echo "== SPECLOSS_RUN_FRAMES=$rf, $1 x $2"; SPECLOSS_RUN_FRAMES=$rf PROF_B=$1 PROF_T=$2 timeout 300 python profiles/time_kernels.py 2>&1 | grep -v default | tee -a $OUT/${T}_ring.txt
  done
done
for rf in 1 0; do
SPECLOSS_RUN_FRAMES=$rf timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${T}_bench_rf$rf.json 2> $OUT/${T}_bench_rf$rf.err; echo "bench rf=$rf rc=$?"; python -c "
import json; d=json.load(open('$OUT/${T}_bench_rf$rf.json')); print(d['ms_per_step'], d['value'], d['roofline']['kernels_ms'], d['roofline']['combine_ms'], d['config']['secondary']['eager_ms_per_step'], d['config']['secondary']['graph_ms_per_step'])"
done
