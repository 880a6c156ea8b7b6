set -u
OUT=gpurun_out; mkdir -p $OUT; T=r3b
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_trainer_step.py -m gpu -x -q -s > $OUT/${T}_tests_2gpu.log 2>&1; echo "tests rc=$?"; tail -8 $OUT/${T}_tests_2gpu.log
timeout 900 python bench.py --gpus 2 --steps 20 --warmup 5 --verbose > $OUT/${T}_bench_n2.json 2> $OUT/${T}_bench_n2.err; echo "bench rc=$?"; cat $OUT/${T}_bench_n2.json; tail -5 $OUT/${T}_bench_n2.err
