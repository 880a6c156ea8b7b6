set -u
OUT=gpurun_out; mkdir -p $OUT; T=r4a
PROF_B=32 PROF_T=192000 timeout 300 python profiles/timeline.py > $OUT/${T}_timeline_32x4s.txt 2>&1; echo "timeline rc=$?"; head -60 $OUT/${T}_timeline_32x4s.txt
PROF_B=16 PROF_T=48000 timeout 300 python profiles/timeline.py > $OUT/${T}_timeline_16x1s.txt 2>&1; echo "timeline rc=$?"; head -50 $OUT/${T}_timeline_16x1s.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${T}_bench_n1.json 2> $OUT/${T}_bench_n1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/${T}_bench_n1.json')); print(d['ms_per_step'], d['value'], d['e2e'])"
