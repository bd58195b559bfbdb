cd $GRAFT_REPO_ROOT
for c in 22 26; do
python bench.py --steps 3 --warmup 3 --no-cpu --eval-cache-log2 $c > gpurun_out/t57_c$c.json 2>> gpurun_out/t57.err
done
tail -3 gpurun_out/t57.err
