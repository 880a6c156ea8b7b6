set -u
OUT=gpurun_out; mkdir -p $OUT; T=r3a
rm -f $OUT/parity_table.tsv
timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${T}_tests.log
timeout 200 python profiles/prof_eager_host.py > $OUT/${T}_host_profile.txt 2>&1; echo "hostprof rc=$?"; head -30 $OUT/${T}_host_profile.txt
timeout 600 python bench.py --steps 20 --warmup 5 --verbose > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"; cat $OUT/${T}_bench.json; tail -5 $OUT/${T}_bench.err
timeout 600 python profiles/trainer_step.py > $OUT/${T}_trainer_step.json 2> $OUT/${T}_trainer_step.err; echo "trainer rc=$?"; cat $OUT/${T}_trainer_step.json; tail -3 $OUT/${T}_trainer_step.err
PROF_B=32 PROF_T=192000 timeout 300 python profiles/time_kernels.py > $OUT/${T}_time_kernels_c4.txt 2>&1; echo "kernels rc=$?"; cat $OUT/${T}_time_kernels_c4.txt
