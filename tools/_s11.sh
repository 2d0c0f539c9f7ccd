set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/s11_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/s11_bench2.json 2> gpurun_out/s11_bench2.err ) 2> gpurun_out/s11_bench2.time
tail -5 gpurun_out/s11_bench2.err
cat gpurun_out/s11_tests.log gpurun_out/s11_bench2.time
