python -m pytest tests/test_ldm_unet_gpu.py tests/test_latent_nets_gpu.py tests/test_ldm_ops_gpu.py -q -x 2>&1 | tail -3
python tools/profile_ldm.py 16 360 640 2>&1 | sed -n 1,34p
python bench.py --config ldm --steps 3 --warmup 3 --no-gpu-baseline > gpurun_out/bench_ldm3.json 2> gpurun_out/bench_ldm3.err; tail -2 gpurun_out/bench_ldm3.err; cat gpurun_out/bench_ldm3.json
