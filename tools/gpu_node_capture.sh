TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
timeout 400 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2_pytest_dist_8gpu.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_dist_8gpu.log; tail -5 gpurun_out/r2_pytest_dist_8gpu.log
timeout 300 $TR bench.py --gpus 8 --sims 800 --games 2048 --vl 4 --window steady --steps 3 --warmup 3 --e2e-games 2048 > gpurun_out/r2_bench_cfg3_vl4_8gpu.json 2> gpurun_out/r2_bench_cfg3_vl4_8gpu.err
timeout 300 $TR -m othellozero_b200.iteration --episodes 4096 --epochs 10 --batch-size 32 > gpurun_out/r2_iteration_8gpu_auto.json 2> gpurun_out/r2_iteration_8gpu_auto.err
tail -c 300 gpurun_out/r2_iteration_8gpu_auto.err
