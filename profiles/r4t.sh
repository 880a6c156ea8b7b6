set -u
OUT=gpurun_out; mkdir -p $OUT; T=r4t
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${T}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${T}_bench_n1.json 2> $OUT/${T}_bench_n1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/${T}_bench_n1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['traffic_capture_is_stale'])"
