timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_t.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_t.log; tail -4 gpurun_out/r2_pytest_t.log
timeout 200 python bench.py --workload tree --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_t_tree_steady.json 2>&1
timeout 200 python bench.py --workload tree --window opening --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_t_tree_opening.json 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/r2_t_selfplay.json 2> gpurun_out/r2_t_selfplay.err
timeout 300 python bench.py --sims 800 --games 512 --vl 8 --window steady --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/r2_t_cfg3_512x8.json 2>&1
timeout 300 python bench.py --sims 800 --games 2048 --vl 4 --window steady --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/r2_t_cfg3_2048x4.json 2>&1
