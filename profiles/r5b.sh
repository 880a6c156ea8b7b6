set -u
OUT=gpurun_out; mkdir -p $OUT; T=r5b
timeout 300 python profiles/time_mel_l1.py 2>&1 | grep -v -i warn | tee $OUT/${T}_time_mel_l1.txt
timeout 900 python -m pytest tests/test_mel_l1.py -m gpu -q > $OUT/${T}_tests_mel_l1.log 2>&1; echo "mel_l1 tests rc=$?"; tail -2 $OUT/${T}_tests_mel_l1.log
