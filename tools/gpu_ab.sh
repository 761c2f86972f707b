# A/B of library builds on one C2 solve loop: bash tools/gpu_ab.sh lib1.so lib2.so ...
for lib in "$@"; do
  echo "== $lib"
  ALTRO_B200_LIB=$lib python tools/gpu_variants.py default 2>&1 | tail -1
done
