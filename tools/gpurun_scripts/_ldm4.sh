python -m pytest tests/test_ldm_unet_gpu.py tests/test_latent_nets_gpu.py -q -x 2>&1 | tail -5
python bench.py --config ldm --steps 3 --warmup 3 > gpurun_out/bench_ldm.json 2> gpurun_out/bench_ldm.err; tail -3 gpurun_out/bench_ldm.err; cat gpurun_out/bench_ldm.json
