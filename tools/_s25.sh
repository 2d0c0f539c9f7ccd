cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_jpeg.py -q -m gpu -x 2>&1 | tail -8
timeout 600 python tools/bench_jpeg.py 256 2>&1 | tail -2 > gpurun_out/s26_bench_jpeg.jsonl; cat gpurun_out/s26_bench_jpeg.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['workload'][:80]); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if 'images_per_s' in k}); print(d['huffman_dev'])"
python tools/_prof_jpeg_host.py 2>&1 | head -3
