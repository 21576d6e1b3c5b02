mkdir -p gpurun_out
for lib in lib_pf1_mc1 lib_pf2_mc1 lib_pf3_mc1 lib_pf1_mc2 lib_pf1_sp1 lib_pf1_sp2 lib_pf1_sp3; do
  echo "== $lib"
  GCN10_CUDA_LIB=$PWD/build/$lib.so timeout 200 python tools/kbench.py --rows-per-cta 8,12,16,24 --profile worldcover --planes 9,18 --tma 1 2>&1 | grep -o '"planes.*'
done
