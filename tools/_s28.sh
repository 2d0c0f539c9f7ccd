cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_jpeg.py -q -m gpu -x 2>&1 | tail -3
python tools/_huff_restart.py
