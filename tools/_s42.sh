cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma2_tile -s 30 -c 8 -o gpurun_out/s42_prof_gemm -f python tools/profile_step.py embed 256 > gpurun_out/s42_prof_gemm.log 2>&1
tail -2 gpurun_out/s42_prof_gemm.log
