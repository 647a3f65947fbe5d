#!/bin/bash
# One line of the cfg5 scaling sweep (run under `gpurun --gpus N`):  bash dev/cfg5_sweep.sh r02v N
tag=$1; n=$2; out=gpurun_out; mkdir -p $out
if [ "$n" = 1 ]; then
  python bench.py --config cfg5 --steps 1 --warmup 1 --e2e-steps 0 > $out/${tag}_bench_cfg5_n1.json 2> $out/${tag}_bench_cfg5_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n --config cfg5 --steps 1 --warmup 1 --e2e-steps 0 > $out/${tag}_bench_cfg5_n$n.json 2> $out/${tag}_bench_cfg5_n$n.err
fi
tail -1 $out/${tag}_bench_cfg5_n$n.json | cut -c1-330; tail -2 $out/${tag}_bench_cfg5_n$n.err | cut -c1-200
