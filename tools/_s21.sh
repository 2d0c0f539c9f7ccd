cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "preprocess" 2>&1 | tail -3
for cfg in "16 75" "16 56" "24 75" "24 100" "32 100" "32 113"; do set -- $cfg; B2C_PRE_TR=$1 B2C_PRE_SMEM_KB=$2 python tools/bench_pre.py 256 512 224 14; done
B2C_PRE_TR=16 B2C_PRE_SMEM_KB=75 python tools/bench_pre.py 512 512 224 32
B2C_PRE_TR=32 python tools/bench_pre.py 512 512 224 32
B2C_PRE_TR=32 python tools/bench_pre.py 64 512 336 14
B2C_PRE_TR=32 python tools/bench_pre.py 32 2048 224 14
