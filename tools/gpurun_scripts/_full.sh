python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; tail -2 gpurun_out/bench_now.err; cat gpurun_out/bench_now.json
