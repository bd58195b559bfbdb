cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_net.py -x -q > gpurun_out/t68_tests.log 2>&1; echo "rc=$?" >> gpurun_out/t68_tests.log
tail -3 gpurun_out/t68_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --eval-cache-log2 0 > gpurun_out/t68_nocache.json 2> gpurun_out/t68.err
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/t68_cache.json 2>> gpurun_out/t68.err
tail -2 gpurun_out/t68.err
