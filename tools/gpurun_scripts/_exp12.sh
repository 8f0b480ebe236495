cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/bench_attn.py 2>&1 | tail -4
echo "== c1"; python bench.py --config c1 --steps 10 --no-cpu-baseline --no-gpu-baseline | cut -c1-220
echo "== c2"; python bench.py --config c2 --steps 5 --no-cpu-baseline --no-gpu-baseline | cut -c1-220
echo "== c2 D3_PDL=1"; D3_PDL=1 python bench.py --config c2 --steps 5 --no-cpu-baseline --no-gpu-baseline | cut -c1-220
echo "== c2 D3_PDL=0"; D3_PDL=0 python bench.py --config c2 --steps 5 --no-cpu-baseline --no-gpu-baseline | cut -c1-220
