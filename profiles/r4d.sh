set -u
OUT=gpurun_out; mkdir -p $OUT; T=r4d
timeout 1200 python -m pytest tests/test_trainer_step.py -m gpu -q -s -k "univnet" > $OUT/${T}_trainer_tests.log 2>&1; echo "tests rc=$?"; grep -E "trainer,|UnivNet|passed|failed|Error|assert" $OUT/${T}_trainer_tests.log | head -40

