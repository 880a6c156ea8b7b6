set -u
OUT=gpurun_out; mkdir -p $OUT; T=r4u
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "target_gradient or trainer_contract or fused_module" > $OUT/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -15 $OUT/${T}_tests.log
