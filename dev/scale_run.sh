#!/bin/bash
# Multi-GPU measurement pass (run under `gpurun --gpus N`): torchrun bench lines, the C-host single-process path with
# its NCCL log, the multi-GPU tests.   bash dev/scale_run.sh r02m 8 "cfg3 cfg4 cfg5"
tag=${1:-rXX}; n=${2:-2}; cfgs=${3:-cfg3}
out=gpurun_out; mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@"; }
timeout 300 python -m pytest tests/test_parity_gpu.py -q -k multi_gpu > $out/${tag}_multi_gpu_test_n$n.log 2>&1; tail -2 $out/${tag}_multi_gpu_test_n$n.log
for cfg in $cfgs; do
  case $cfg in
    cfg5) steps=1; warm=1;;
    cfg4) steps=3; warm=3;;
    *)    steps=20; warm=3;;
  esac
  run --config $cfg --steps $steps --warmup $warm > $out/${tag}_bench_${cfg}_n$n.json 2> $out/${tag}_bench_${cfg}_n$n.err
  tail -1 $out/${tag}_bench_${cfg}_n$n.json | cut -c1-300
  tail -3 $out/${tag}_bench_${cfg}_n$n.err | cut -c1-300
done
# the reference's C symbol from one process, HVB200_NGPUS = 1..n (no torch), with NCCL's own log
for cfg in $cfgs; do
  [ $cfg = cfg5 ] && continue          # the drop-in symbols stop at d = 31 like the reference
  reps=5; ks=""; [ $cfg = cfg4 ] && { reps=2; ks="--ngpus $n"; }       # cfg4 on one GPU is 10 s per call: only the N-GPU row
  NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$out/${tag}_chost_${cfg}_nccl.%h.%p.log python dev/multi_gpu_c_host.py --config $cfg --reps $reps $ks \
      --json $out/${tag}_chost_${cfg}_n$n.json > $out/${tag}_chost_${cfg}_n$n.log 2>&1
  tail -4 $out/${tag}_chost_${cfg}_n$n.log | cut -c1-400
done
