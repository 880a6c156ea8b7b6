set -u
OUT=gpurun_out; mkdir -p $OUT; T=${1:-r3f}
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${T}_tests.log
for eo in 0 1; do
  echo "== SPECLOSS_EO_2048=$eo, 16 x 1 s"; SPECLOSS_EO_2048=$eo timeout 300 python profiles/time_kernels.py 2>&1 | tee $OUT/${T}_time_kernels_eo$eo.txt
  echo "== SPECLOSS_EO_2048=$eo, 32 x 4 s"; SPECLOSS_EO_2048=$eo PROF_B=32 PROF_T=192000 timeout 300 python profiles/time_kernels.py 2>&1 | tee $OUT/${T}_time_kernels_c4_eo$eo.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/${T}_bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['kernels_ms'], d['roofline']['combine_ms'], d['config']['secondary'])"
