set -u
N=${1:?n}
OUT=gpurun_out; mkdir -p $OUT; T=r5e
nvidia-smi -L | head -8
timeout 900 python bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${T}_bench_n$N.json 2> $OUT/${T}_bench_n$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/${T}_bench_n$N.json"))
print("N", d["n_gpus"], "ms", d["ms_per_step"], "value", d["value"], "e2e", {k: v for k, v in d["e2e"].items() if k != "timing"})
print("parity", json.dumps(d["sharded_parity"])[:1200])
print("secondary", d["config"]["secondary"])
print("parallelism", d["config"]["parallelism"])
PY
tail -3 $OUT/${T}_bench_n$N.err
