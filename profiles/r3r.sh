set -u
OUT=gpurun_out; mkdir -p $OUT; T=r3r
for it in 6 8 12 16 24 32; do
  echo "== SPECLOSS_MEL_ITER=$it (32 x 4 s)"
  SPECLOSS_MEL_ITER=$it PROF_B=32 PROF_T=192000 timeout 300 python profiles/time_kernels.py 2>&1 | grep "mel2048" | tee -a $OUT/${T}_mel_iter.txt
done
