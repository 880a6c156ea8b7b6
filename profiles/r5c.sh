set -u
OUT=gpurun_out; mkdir -p $OUT; T=r5c
export PROF_B=32 PROF_T=192000 PROF_STEPS=3
timeout 120 python profiles/prof_step.py > $OUT/${T}_plain.log 2>&1 || { echo "plain run failed"; cat $OUT/${T}_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $OUT/${T}_launches_32x4s.csv python profiles/prof_step.py > $OUT/${T}_ncu_launches.log 2>&1; echo "launches rc=$?"
PROF_STEPS=2 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"transform|combine" -s 6 -c 6 -o $OUT/${T}_full_32x4s -f python profiles/prof_step.py > $OUT/${T}_ncu_full.log 2>&1; echo "full rc=$?"
ncu -i $OUT/${T}_full_32x4s.ncu-rep --page raw --csv > $OUT/${T}_full_32x4s_raw.csv 2>/dev/null
export PROF_B=256
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $OUT/${T}_launches_256x4s.csv python profiles/prof_step.py > $OUT/${T}_ncu_launches_c3.log 2>&1; echo "launches c3 rc=$?"
export PROF_B=16 PROF_T=48000
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $OUT/${T}_launches_16x1s.csv python profiles/prof_step.py > $OUT/${T}_ncu_launches_c1.log 2>&1; echo "launches c1 rc=$?"
