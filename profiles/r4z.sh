set -u
OUT=gpurun_out; mkdir -p $OUT; T=r4z
for it in 16 0; do
  echo "== SPECLOSS_MEL_ITER=$it"
  SPECLOSS_MEL_ITER=$it timeout 300 python profiles/time_kernels.py 2>&1 | grep -E "mel2048|all4" | tee -a $OUT/${T}_mel_schedule.txt
  SPECLOSS_MEL_ITER=$it PROF_B=32 PROF_T=192000 timeout 300 python profiles/time_kernels.py 2>&1 | grep -E "mel2048|all4" | tee -a $OUT/${T}_mel_schedule.txt
  SPECLOSS_MEL_ITER=$it PROF_B=256 PROF_T=192000 timeout 300 python profiles/time_kernels.py 2>&1 | grep -E "mel2048|all4" | tee -a $OUT/${T}_mel_schedule.txt
done
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $OUT/${T}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${T}_bench_n1.json 2> $OUT/${T}_bench_n1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/${T}_bench_n1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernels_ms'], d['roofline']['traffic_capture_is_stale'])"
