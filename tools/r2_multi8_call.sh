#!/bin/bash
# Eight-GPU call: MuS bench at N = 8 (in-bench parity, halo overlap on / off), the per-operation timeline of one step on rank 0,
# REMuS 1M at N = 8, and configs[4]'s shape (REMuS, 4M nodes, hidden 256) for a few steps on the fp32 kernels.
mkdir -p gpurun_out
N=${NGPU:-8}
run() { name=$1; shift
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 \
      bench.py --gpus $N "$@" > gpurun_out/r2j_${name}_n$N.json 2> gpurun_out/r2j_${name}_n$N.err; }
run mus --steps 20 --warmup 5
if [ -z "$SKIP_EXTRA" ]; then
run mus_overlap --steps 20 --warmup 5 --overlap --skip-parity
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633 \
    tools/partition_timeline.py > gpurun_out/r2j_timeline_n$N.txt 2> gpurun_out/r2j_timeline_n$N.err
fi
run remus --model remus --steps 10 --warmup 3
if [ "$N" = "8" ] && [ -z "$SKIP_C4" ]; then
  run remus_4m_h256 --model remus --nodes 4000000 --hidden 256 --steps 4 --warmup 1 --skip-parity
fi
for f in gpurun_out/r2j_*_n$N.json; do echo $f; head -c 600 $f; echo; done; for f in gpurun_out/r2j_*_n$N.err; do tail -n 3 $f; done
