cd $GRAFT_REPO_ROOT
python -m pytest tests/test_pipeline_gpu.py tests/test_unet_gpu.py -x -q 2>&1 | tail -6
echo "== bench c1"; python bench.py --config c1 --steps 10 --no-cpu-baseline --no-gpu-baseline | cut -c1-330
echo "== bench c2"; python bench.py --config c2 --steps 5 --no-cpu-baseline --no-gpu-baseline | cut -c1-330
