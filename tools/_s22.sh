cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resample_kernel -s 3 -c 1 -o gpurun_out/s22_pre -f python tools/bench_pre.py 128 512 224 14 > gpurun_out/s22_pre.log 2>&1
tail -3 gpurun_out/s22_pre.log
