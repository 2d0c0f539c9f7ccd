#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list, ncu --set full captures.  Outputs -> gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag] [what...]'
#   what: tests smoke bench refbench multi launches full gemmcmp archs attntest attnbench pre huff train imgstats similar jpeg pipeline lanes power
set -u
TAG=${1:-run}
shift || true
WHAT=${*:-tests bench launches full}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_smi.txt 2>&1
for w in $WHAT; do
  case $w in
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1
      echo "tests exit $?" | tee -a $OUT/${TAG}_tests.log; tail -5 $OUT/${TAG}_tests.log ;;
    smoke)
      timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; tail -3 $OUT/${TAG}_smoke.log ;;
    bench)
      timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
      echo "bench exit $?"; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err ;;
    multi)
      N=$(nvidia-smi -L | wc -l)
      timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/${TAG}_multi_tests.log 2>&1; tail -3 $OUT/${TAG}_multi_tests.log
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
        bench.py --gpus $N --steps 6 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
      echo "bench N=$N exit $?"; cat $OUT/${TAG}_bench_n$N.json; tail -3 $OUT/${TAG}_bench_n$N.err ;;
    gemmcmp)
      timeout 300 python tools/gemm_compare.py 1024 > $OUT/${TAG}_gemm_compare.json 2> $OUT/${TAG}_gemm_compare.err; cat $OUT/${TAG}_gemm_compare.json ;;
    archs)
      : > $OUT/${TAG}_archs.jsonl
      timeout 300 python tools/bench_arch.py ViT-H-14/laion2b_s32b_b79k --fc --batch 128 >> $OUT/${TAG}_archs.jsonl 2>> $OUT/${TAG}_archs.err
      timeout 300 python tools/bench_arch.py ViT-L-14-336/openai --batch 64 >> $OUT/${TAG}_archs.jsonl 2>> $OUT/${TAG}_archs.err
      timeout 300 python tools/bench_arch.py ViT-B-32/openai --batch 512 >> $OUT/${TAG}_archs.jsonl 2>> $OUT/${TAG}_archs.err
      cat $OUT/${TAG}_archs.jsonl; tail -3 $OUT/${TAG}_archs.err ;;
    attntest)
      timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "attention or encode_image" > $OUT/${TAG}_attntest.log 2>&1; tail -15 $OUT/${TAG}_attntest.log ;;
    attnbench)
      timeout 120 python tools/bench_attn.py 2>&1 | tee $OUT/${TAG}_attnbench.jsonl
      timeout 120 python tools/bench_attn.py 256 577 16 64 2>&1 | tee -a $OUT/${TAG}_attnbench.jsonl ;;
    pre)
      timeout 300 python tools/bench_pre.py 256 512 224 14 | tee $OUT/${TAG}_bench_pre.jsonl
      timeout 300 python tools/bench_pre.py 32 2048 224 14 | tee -a $OUT/${TAG}_bench_pre.jsonl ;;
    huff)
      timeout 300 python tools/bench_huff_restart.py | tee $OUT/${TAG}_bench_huff.txt ;;
    train)
      timeout 600 python -m pytest tests/test_train.py -m gpu -x -q > $OUT/${TAG}_train_tests.log 2>&1; tail -15 $OUT/${TAG}_train_tests.log
      timeout 300 python tools/bench_train.py > $OUT/${TAG}_bench_train.json 2> $OUT/${TAG}_bench_train.err; cat $OUT/${TAG}_bench_train.json; tail -3 $OUT/${TAG}_bench_train.err ;;
    imgstats)
      timeout 600 python -m pytest tests/test_gpu_imgstats.py -m gpu -x -q > $OUT/${TAG}_imgstats_tests.log 2>&1; tail -15 $OUT/${TAG}_imgstats_tests.log
      timeout 300 python tools/bench_imgstats.py > $OUT/${TAG}_bench_imgstats.json 2> $OUT/${TAG}_bench_imgstats.err; cat $OUT/${TAG}_bench_imgstats.json; tail -3 $OUT/${TAG}_bench_imgstats.err ;;
    similar)
      timeout 600 python -m pytest tests/test_gpu_similar.py -m gpu -x -q > $OUT/${TAG}_similar_tests.log 2>&1; tail -15 $OUT/${TAG}_similar_tests.log
      timeout 300 python tools/bench_similar.py > $OUT/${TAG}_bench_similar.json 2> $OUT/${TAG}_bench_similar.err; cat $OUT/${TAG}_bench_similar.json; tail -3 $OUT/${TAG}_bench_similar.err ;;
    jpeg)
      timeout 600 python -m pytest tests/test_jpeg.py -x -q > $OUT/${TAG}_jpeg_tests.log 2>&1; tail -3 $OUT/${TAG}_jpeg_tests.log
      timeout 300 python tools/bench_jpeg.py 256 > $OUT/${TAG}_bench_jpeg.json 2> $OUT/${TAG}_bench_jpeg.err; cat $OUT/${TAG}_bench_jpeg.json ;;
    pipeline)
      B2C_DRIVER_TIMING=1 timeout 900 python tools/bench_pipeline.py --n 8192 2> $OUT/${TAG}_pipeline.err | tee $OUT/${TAG}_pipeline.jsonl
      grep main-thread $OUT/${TAG}_pipeline.err | tail -1 ;;
    lanes)
      timeout 400 python tools/bench_lanes.py --batches 256 --lanes 1,2,4 --fused 0,1 2> $OUT/${TAG}_lanes.err | tee $OUT/${TAG}_lanes.jsonl ;;
    power)
      timeout 500 python tools/power_probe.py 512 > $OUT/${TAG}_power.json 2> $OUT/${TAG}_power.err; cat $OUT/${TAG}_power.json ;;
    refbench)
      timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_refbench.json 2> $OUT/${TAG}_refbench.err
      cat $OUT/${TAG}_refbench.json ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
        --log-file $OUT/${TAG}_launches_embed.csv python tools/profile_step.py embed 256 > $OUT/${TAG}_launches_embed.log 2>&1
      timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file $OUT/${TAG}_launches_dedup.csv python tools/profile_step.py dedup 200000 > $OUT/${TAG}_launches_dedup.log 2>&1
      echo "launches done" ;;
    full)
      # one capture per kernel family: a few launches each (ncu replays ~40x per kernel)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma2_tile -s 30 -c 8 \
        -o $OUT/${TAG}_prof_gemm -f python tools/profile_step.py embed 256 > $OUT/${TAG}_prof_gemm.log 2>&1
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 \
        -o $OUT/${TAG}_prof_attn -f python tools/profile_attn.py 512 > $OUT/${TAG}_prof_attn.log 2>&1
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:layernorm -c 1 \
        -o $OUT/${TAG}_prof_ln -f python tools/profile_step.py embed 128 > $OUT/${TAG}_prof_ln.log 2>&1
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:resample_kernel -c 1 \
        -o $OUT/${TAG}_prof_pre -f python tools/profile_step.py embed 128 > $OUT/${TAG}_prof_pre.log 2>&1
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:"umma|dedup" -s 1 -c 2 \
        -o $OUT/${TAG}_prof_dedup -f python tools/profile_step.py dedup 100000 > $OUT/${TAG}_prof_dedup.log 2>&1
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:jpeg_huff_kernel -s 2 -c 1 \
        -o $OUT/${TAG}_prof_huff -f python tools/bench_huff_restart.py > $OUT/${TAG}_prof_huff.log 2>&1
      ls -la $OUT | tail -20 ;;
  esac
done
