# 1 -> 8 GPU weak scaling of the default bench line on one node (final binary)
for n in 8 4 2; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_scale_$n.json 2> gpurun_out/r2_scale_$n.err
done
timeout 240 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/r2_scale_1.json 2> gpurun_out/r2_scale_1.err
