set -u
OUT=gpurun_out; mkdir -p $OUT; T=r3p
for caps in "0 0" "8 8" "6 8" "6 6" "4 6" "8 12" "6 10" "10 12"; do set -- $caps
  echo "== SPECLOSS_WARPS_2048=$1 SPECLOSS_WARPS_SMALL=$2 (16 x 1 s)"
  SPECLOSS_WARPS_2048=$1 SPECLOSS_WARPS_SMALL=$2 timeout 300 python profiles/time_kernels.py 2>&1 | grep -v default | tee -a $OUT/${T}_coresidency.txt
done
