#!/bin/bash
# One multi-GPU box call of a round: H2D ceilings, the config-4 sweep with the NCCL prediction gather, the forward bench and the
# config-5 training bench at N GPUs.  Usage (under gpurun --gpus N):  bash tools/run_multigpu.sh N
N=${1:-8}
OUT=gpurun_out/multigpu_n$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
python tools/h2d_bw_multi.py --sets 0 0,1 0,1,2,3 0,2,4,6 0,1,2,3,4,5,6,7 > $OUT/h2d_bw.jsonl 2> $OUT/h2d_bw.err
$TR --master-port 29601 tools/sweep.py --videos $((1024 * N)) --batch 64 --pool 64 > $OUT/sweep.json 2> $OUT/sweep.err
$TR --master-port 29602 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
$TR --master-port 29603 bench.py --train --gpus $N --steps 6 --warmup 3 > $OUT/train.json 2> $OUT/train.err
tail -c 600 $OUT/sweep.json; echo; tail -c 300 $OUT/bench.json; echo; tail -c 900 $OUT/train.json; echo; cat $OUT/h2d_bw.jsonl
