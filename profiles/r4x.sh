set -u
OUT=gpurun_out; mkdir -p $OUT; T=r4x
D=dl_speech_enhancement_b200
cp $D/libspecloss.so /tmp/lib_default.so
for v in default mb3 mb4; do
  if [ $v = default ]; then cp /tmp/lib_default.so $D/libspecloss.so; else cp $D/variant_$v.so $D/libspecloss.so; fi
  echo "== $v"; timeout 300 python profiles/time_mel_l1.py 2>&1 | grep -v -i warn | tee $OUT/${T}_time_mel_l1_$v.txt
done
cp /tmp/lib_default.so $D/libspecloss.so
