cd $GRAFT_REPO_ROOT
python tools/bench_attn.py 256 577 16 64 2>&1 | tail -2
B2C_ATTN_T577=3 python tools/bench_attn.py 256 577 16 64 2>&1 | tail -2
python tools/bench_attn.py 37 577 16 64 2>&1 | tail -1
python tools/bench_attn.py 64 385 12 64 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention or encode_image_vs_oracle" 2>&1 | tail -3
