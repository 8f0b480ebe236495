cd $GRAFT_REPO_ROOT
python -m pytest tests/test_pipeline_gpu.py -x -q -k whole_loop 2>&1 | tail -40
python -m pytest tests/test_frontend_gpu.py -x -q 2>&1 | tail -15
