set -u
OUT=gpurun_out; mkdir -p $OUT; T=r5a
nvidia-smi -L | wc -l
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${T}_tests_2gpu_box.log 2>&1; echo "tests rc=$?"; tail -2 $OUT/${T}_tests_2gpu_box.log
