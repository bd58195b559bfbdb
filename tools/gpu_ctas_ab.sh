for lib in liboz_b200.so liboz_b200_ctas8.so liboz_b200_ctas6.so; do
  export OZ_B200_LIB=$PWD/othellozero_b200/$lib
  timeout 200 python bench.py --workload tree --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_ab_tree_steady_$lib.json 2>&1
  timeout 200 python bench.py --workload tree --window opening --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_ab_tree_opening_$lib.json 2>&1
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/r2_ab_sp_$lib.json 2>&1
done
