python -m pytest tests/test_ldm_unet_gpu.py tests/test_latent_nets_gpu.py tests/test_ldm_ops_gpu.py -q -x 2>&1 | tail -4
python bench.py --config ldm --steps 3 --warmup 3 --no-gpu-baseline > gpurun_out/bench_ldm2.json 2> gpurun_out/bench_ldm2.err; tail -3 gpurun_out/bench_ldm2.err; cat gpurun_out/bench_ldm2.json
