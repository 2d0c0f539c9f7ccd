cd $GRAFT_REPO_ROOT
B2C_DRIVER_TIMING=1 timeout 1200 python tools/bench_pipeline.py --n 8192 2> gpurun_out/s27_pipeline.err | tee gpurun_out/s27_pipeline.jsonl
grep main-thread gpurun_out/s27_pipeline.err | tail -4
tail -3 gpurun_out/s27_pipeline.err
