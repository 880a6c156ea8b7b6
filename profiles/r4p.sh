set -u
OUT=gpurun_out; mkdir -p $OUT; T=r4p
for i in 1 2 3; do timeout 600 python -m pytest tests/test_trainer_step.py -m gpu -q -s > $OUT/${T}_trainer_tests_$i.log 2>&1; echo "trainer tests run $i rc=$?"; grep -E "^E   +Assert|passed|failed" $OUT/${T}_trainer_tests_$i.log | head -5; done
grep -E "trainer," $OUT/${T}_trainer_tests_1.log | cut -c1-420
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $OUT/${T}_tests.log
timeout 300 python profiles/time_kernels.py 2>&1 | tee $OUT/${T}_time_kernels.txt
PROF_B=32 PROF_T=192000 timeout 300 python profiles/time_kernels.py 2>&1 | tee $OUT/${T}_time_kernels_c4.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${T}_bench_n1.json 2> $OUT/${T}_bench_n1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/${T}_bench_n1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernels_ms'], d['config']['secondary'])"
