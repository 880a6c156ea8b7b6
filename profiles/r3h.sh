set -u
OUT=gpurun_out; mkdir -p $OUT; T=r3h
for v in ${VARIANTS:-w12uq}; do
  for sz in "16 48000" "32 192000"; do set -- $sz
    echo "== variant $v, $1 x $2"; PROF_LIB=profiles/variants/libspecloss_$v.so PROF_B=$1 PROF_T=$2 timeout 300 python profiles/time_kernels.py 2>&1 | grep "2048" | tee -a $OUT/${T}_variants.txt
  done
done
