set -u
OUT=gpurun_out; mkdir -p $OUT; T=r3n
rm -f $OUT/parity_table.tsv
timeout 900 python -m pytest tests -m gpu -q -s > $OUT/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${T}_tests.log; grep -n "trainer tensors\|free-running" $OUT/${T}_tests.log | cut -c1-330
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/${T}_bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['kernels_ms'], d['roofline']['combine_ms'], d['config']['secondary'])"
