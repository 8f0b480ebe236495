cd $GRAFT_REPO_ROOT
python -m pytest tests/test_unet_gpu.py tests/test_pipeline_gpu.py tests/test_kernels_gpu.py tests/test_guided_gpu.py -x -q 2>&1 | tail -5
for pdl in 1 0; do
echo "== D3_PDL=$pdl c1"; D3_PDL=$pdl python bench.py --config c1 --steps 10 --no-cpu-baseline --no-gpu-baseline | cut -c1-220
echo "== D3_PDL=$pdl c2"; D3_PDL=$pdl python bench.py --config c2 --steps 5 --no-cpu-baseline --no-gpu-baseline | cut -c1-220
done
