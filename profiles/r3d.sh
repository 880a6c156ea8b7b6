set -u
OUT=gpurun_out; mkdir -p $OUT; T=r3d
timeout 300 python profiles/sanitize_step.py > $OUT/${T}_sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -3 $OUT/${T}_sanitize_plain.log
for tool in memcheck racecheck initcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_step.py > $OUT/${T}_sanitizer_${tool}.log 2>&1; echo "$tool rc=$?"; grep -n "ERROR SUMMARY\|RACECHECK SUMMARY\|sanitize_step done\|Error\|hazard" $OUT/${T}_sanitizer_${tool}.log | head -8
done
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${T}_tests.log
