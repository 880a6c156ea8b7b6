set -u
OUT=gpurun_out; mkdir -p $OUT; T=r4k
for i in 1 2 3 4; do timeout 600 python -m pytest tests/test_trainer_step.py -m gpu -q -s > $OUT/${T}_trainer_tests_$i.log 2>&1; echo "trainer tests run $i rc=$?"; grep -E "^E   +Assert|passed|failed" $OUT/${T}_trainer_tests_$i.log | head -5; done
grep -E "trainer," $OUT/${T}_trainer_tests_1.log
