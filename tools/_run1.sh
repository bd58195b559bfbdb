cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t72_tests.log 2>&1; echo "rc=$?" >> gpurun_out/t72_tests.log
tail -3 gpurun_out/t72_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t72_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/t72_smoke.log; tail -1 gpurun_out/t72_smoke.log
python bench.py > gpurun_out/t72_bench.json 2> gpurun_out/t72.err; echo "bench rc=$?" >> gpurun_out/t72.err
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --eval-cache-log2 0 > gpurun_out/t72_nocache.json 2>> gpurun_out/t72.err
python bench.py --steps 10 --warmup 3 --no-cpu --conv3 wino > gpurun_out/t72_wino.json 2>> gpurun_out/t72.err
ncu --set full --clock-control none --import-source on -k regex:"conv2_table_gather|oz_gemm" -s 90 -c 6 -o gpurun_out/prof_final2_r1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --eval-cache-log2 0 > gpurun_out/t72_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 420 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/t72_ncu2.log 2>&1
tail -2 gpurun_out/t72.err
