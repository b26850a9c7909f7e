#!/bin/bash
# Multi-GPU measurement call of the next round:   gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_scaling_call.sh'
# (charged 8x the box time: about 2 minutes of wall clock per line below).  MuS-GNN strong scaling on the 1M-node mesh, then the first
# timing of the REMuS-GNN edge-halo partition.  EDGE_MODE=m selects an experimental edge-kernel variant for every run.
mkdir -p gpurun_out
M=${EDGE_MODE:-0}
run() {   # name, gpus, extra args
  local name=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --skip-cpu-baseline --edge-mode $M "$@" > gpurun_out/r2_${name}_n1.json 2> gpurun_out/r2_${name}_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 --edge-mode $M "$@" > gpurun_out/r2_${name}_n${n}.json 2> gpurun_out/r2_${name}_n${n}.err
  fi
  echo "$name n=$n exit $?: $(head -c 300 gpurun_out/r2_${name}_n${n}.json)"
}
for n in 1 2 4 8; do run mus $n; done
for n in 1 2 4 8; do run remus $n --model remus; done
