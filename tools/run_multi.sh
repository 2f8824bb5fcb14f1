#!/bin/bash
# usage: tools/run_multi.sh N  — the multi-GPU lines of BASELINE configs 1-5 on N GPUs of one box (torchrun, one rank per GPU)
N=$1; OUT=gpurun_out; P=29500
run() { name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N "$@" > $OUT/r02_${name}_${N}gpu.json 2> $OUT/r02_${name}_${N}gpu.err || tail -c 600 $OUT/r02_${name}_${N}gpu.err; P=$((P+1)); cut -c1-260 $OUT/r02_${name}_${N}gpu.json; }
run kitti --steps 10 --warmup 3
run nuscenes --workload nuscenes --steps 10 --warmup 3
run lokitti_kpfcnn --workload lokitti --net kpfcnn --steps 6 --warmup 3
if [ "$N" = "8" ]; then
  run lokitti_kpfcnn_64pairs --workload lokitti --net kpfcnn --streams 1 --batch 8 --steps 6 --warmup 3
  run train --mode train --steps 30 --warmup 5 --pairs 2
fi
