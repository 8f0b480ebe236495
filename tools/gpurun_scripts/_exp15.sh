cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== c1"; python bench.py --config c1 --steps 10 --no-cpu-baseline --no-gpu-baseline | cut -c1-220
echo "== c2"; python bench.py --config c2 --steps 5 --no-cpu-baseline --no-gpu-baseline | cut -c1-220
