cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for c in c2 c3 c5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --config $c --steps 3 --warmup 3 --no-gpu-baseline > gpurun_out/bench_${c}_8gpu.json 2> gpurun_out/bench_${c}_8gpu.err
  echo "== $c rc=$?"; cut -c1-400 gpurun_out/bench_${c}_8gpu.json; tail -3 gpurun_out/bench_${c}_8gpu.err
done
