cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t46_tests.log 2>&1; echo "rc=$?" >> gpurun_out/t46_tests.log
tail -4 gpurun_out/t46_tests.log
python bench.py --workload perft --steps 10 --warmup 3 > gpurun_out/t46_perft.json 2> gpurun_out/t46.err
cat gpurun_out/t46_perft.json | head -c 2500
