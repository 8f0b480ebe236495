cd $GRAFT_REPO_ROOT
echo "== bench c2 (default)"; python bench.py --steps 5 > gpurun_out/bench_c2.json; cat gpurun_out/bench_c2.json
echo "== reference arm"; python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json; cat gpurun_out/bench_ref.json
echo "== per-layer profile"; python tools/profile_layers.py 16 fp16 > gpurun_out/layers_b16.txt; cat gpurun_out/layers_b16.txt
echo "== identity slab off"; D3_UNET_IDENTITY_SLAB=0 python tools/profile_layers.py 16 fp16 | head -14
echo "== L2 chain"; python tools/exp_l2_chain.py
echo "== bench c1"; python bench.py --config c1 --steps 5 --no-cpu-baseline > gpurun_out/bench_c1.json; cat gpurun_out/bench_c1.json
echo "== bench c3"; python bench.py --config c3 --steps 5 --no-cpu-baseline > gpurun_out/bench_c3.json; cat gpurun_out/bench_c3.json
echo "== bench c5"; python bench.py --config c5 --steps 3 --no-cpu-baseline > gpurun_out/bench_c5.json; cat gpurun_out/bench_c5.json
