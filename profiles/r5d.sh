set -u
OUT=gpurun_out; mkdir -p $OUT; T=r5d
timeout 1500 python -m pytest tests/ -x -q -m gpu > $OUT/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $OUT/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${T}_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/${T}_bench_reference.json 2> $OUT/${T}_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${T}_bench_n1.json 2> $OUT/${T}_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/${T}_bench_n1.json")); r=json.load(open("$OUT/${T}_bench_reference.json"))
print("ours", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "stale", d["roofline"]["traffic_capture_is_stale"], "frac", d["roofline"]["frac"], d["roofline"]["fp32_frac"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("ref", r["value"], r["ms_per_step"], r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"])
print("one-call", d["config"]["one_call_extension"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
PY
