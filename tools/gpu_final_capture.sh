# Final evidence of the round (one B200): tests, the default bench line, the CPU arm, launch list, ncu of one engine step, ncu of the tree-only job
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_final.log; tail -4 gpurun_out/r2_pytest_final.log
timeout 400 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_ref.json 2> gpurun_out/r2_bench_final_ref.err
timeout 200 python bench.py --workload tree --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_bench_tree_steady.json 2>&1
timeout 200 python bench.py --workload tree --window opening --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_bench_tree_opening.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 420 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tree_step|conv2_table|oz_gemm" -s 2100 -c 8 -f -o gpurun_out/r2_prof_step2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2_ncu_step2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tree_step -s 0 -c 1 -f -o gpurun_out/r2_prof_tree2 python bench.py --workload tree --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_ncu_tree2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
