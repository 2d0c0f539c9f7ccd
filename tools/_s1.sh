set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention or layernorm or transformers or encode_image_vs_oracle or lanes" 2>&1 | tail -30 > gpurun_out/s1_tests.log
timeout 600 python tools/bench_attn.py --vars 0,1,3,5,7,11,15 > gpurun_out/s1_attn_ab.jsonl 2> gpurun_out/s1_attn_ab.err
timeout 300 python tools/bench_attn.py --vars 0,3,7 512 >> gpurun_out/s1_attn_ab.jsonl 2>> gpurun_out/s1_attn_ab.err
timeout 600 python bench.py --no-variants --dedup-n 0 --no-cpu-baseline > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
cat gpurun_out/s1_tests.log gpurun_out/s1_attn_ab.jsonl; tail -3 gpurun_out/s1_attn_ab.err; cut -c1-600 gpurun_out/s1_bench.json
