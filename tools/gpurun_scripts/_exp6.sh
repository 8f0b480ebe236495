cd $GRAFT_REPO_ROOT
python tools/exp_two_streams.py 16
python tools/exp_two_streams.py 32
