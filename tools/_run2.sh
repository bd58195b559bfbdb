cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/t49_bench2.json 2> gpurun_out/t49.err
tail -c 1500 gpurun_out/t49_bench2.json; tail -3 gpurun_out/t49.err
