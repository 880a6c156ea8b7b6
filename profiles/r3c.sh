set -u
OUT=gpurun_out; mkdir -p $OUT; T=r3c
timeout 900 python -m pytest tests/test_mel_l1.py tests/test_trainer_step.py -m gpu -q -s > $OUT/${T}_tests.log 2>&1; echo "tests rc=$?"; grep -n "trainer\|passed\|failed\|Error" $OUT/${T}_tests.log | tail -15
timeout 300 python profiles/time_mel_l1.py > $OUT/${T}_time_mel_l1.txt 2>&1; echo "mel_l1 rc=$?"; cat $OUT/${T}_time_mel_l1.txt
