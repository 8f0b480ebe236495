cd $GRAFT_REPO_ROOT
python -m pytest tests/test_conv_gpu.py tests/test_kernels_gpu.py -x -q 2>&1 | tail -15
echo "== EPI=1 (stmatrix)"; python tools/bench_sustained.py 2 0 1
echo "== EPI=0 (lane exchange)"; D3_CONV_EPI=0 python tools/bench_sustained.py 2 0 1
python -m pytest tests/test_unet_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -15
echo "== fp16 bench"; python bench.py --dtype fp16 --steps 5 --no-cpu-baseline
