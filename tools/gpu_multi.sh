#!/bin/bash
# One multi-GPU box:  tools/gpu_multi.sh N TAG  ->  gpurun_out/TAG_n{N}_*: the distinct-GPU tests, bench.py for every BASELINE
# config at N ranks (torchrun, as the driver launches it) and the host<->device copy matrix of the box
N=$1; TAG=${2:-r02}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_n${N}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_one_shot.py -m gpu -x -q > $OUT/${TAG}_n${N}_pytest.log 2>&1; tail -2 $OUT/${TAG}_n${N}_pytest.log
for WL in C3 C5 C4; do
  STEPS=10; [ $WL = C4 ] && STEPS=5
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 3 --workload $WL \
      > $OUT/${TAG}_n${N}_bench_${WL}.json 2> $OUT/${TAG}_n${N}_bench_${WL}.err
  tail -c 400 $OUT/${TAG}_n${N}_bench_${WL}.json; echo
done
if [ $N -ge 8 ]; then  # the N = 4 headline figure on the same box (GPUs 0-3)
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --workload C3 \
      > $OUT/${TAG}_n4on8_bench_C3.json 2> $OUT/${TAG}_n4on8_bench_C3.err
  tail -c 300 $OUT/${TAG}_n4on8_bench_C3.json; echo
fi
SETS="0"; [ $N -ge 2 ] && SETS="$SETS 0,1"; [ $N -ge 4 ] && SETS="$SETS 0,1,2,3"; [ $N -ge 8 ] && SETS="$SETS 4,5,6,7 0,1,2,3,4,5,6,7 0,4 0,2,4,6"
timeout 600 aruco3_b200/csrc/build/pcie_matrix --mb 1024 --passes 4 $SETS > $OUT/${TAG}_n${N}_pcie_matrix.jsonl 2> $OUT/${TAG}_n${N}_pcie_matrix.err
tail -3 $OUT/${TAG}_n${N}_pcie_matrix.jsonl
