set -u
OUT=gpurun_out; mkdir -p $OUT; T=r4r
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${T}_bench_reference.json 2> $OUT/${T}_bench_reference.err; echo "ref rc=$?"; cut -c1-700 $OUT/${T}_bench_reference.json
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${T}_bench_n1.json 2> $OUT/${T}_bench_n1.err; echo "bench rc=$?"; cat $OUT/${T}_bench_n1.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${T}_smoke.log
nproc; free -g | head -2
