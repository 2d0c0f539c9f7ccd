set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/s10_tests.log
( time timeout 900 python bench.py > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err ) 2> gpurun_out/s10_bench.time
tail -5 gpurun_out/s10_bench.err
( time timeout 900 python bench.py --impl reference > gpurun_out/s10_ref.json 2> gpurun_out/s10_ref.err ) 2> gpurun_out/s10_ref.time
cat gpurun_out/s10_tests.log gpurun_out/s10_bench.time gpurun_out/s10_ref.time
