set -u
OUT=gpurun_out; mkdir -p $OUT; T=r4o
for cfg in "SPECLOSS_EO_2048=1" "SPECLOSS_EO_2048=0" "SPECLOSS_RUN_FRAMES=8" "SPECLOSS_RUN_FRAMES=3 SPECLOSS_EO_2048=1" "SPECLOSS_SERIAL=1"; do
  env $cfg timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $OUT/${T}_tests_$(echo $cfg | tr ' =' '__').log 2>&1; echo "$cfg rc=$?"; tail -1 $OUT/${T}_tests_$(echo $cfg | tr ' =' '__').log
done
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -q -k "sharded or two_gpu" > $OUT/${T}_tests_2gpu.log 2>&1; echo "2gpu rc=$?"; tail -2 $OUT/${T}_tests_2gpu.log
