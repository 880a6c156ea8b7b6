set -u
OUT=gpurun_out; mkdir -p $OUT; T=r3i
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${T}_tests.log
bash profiles/r3h.sh
timeout 200 python profiles/prof_eager_host.py > $OUT/${T}_host_profile.txt 2>&1; head -12 $OUT/${T}_host_profile.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/${T}_bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['kernels_ms'], d['roofline']['combine_ms'], d['config']['secondary'])"
