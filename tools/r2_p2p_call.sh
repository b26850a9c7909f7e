#!/bin/bash
# N-GPU call (NGPU, default 2): peer-memory halo (g4c_halo_put) against the NCCL halo -- parity tests, per-operation timeline,
# bench lines with the in-bench parity check.
N=${NGPU:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs -s -k partitioned 2>&1 | grep -E "world=|passed|failed|skipped|Error|error" | tail -30 > gpurun_out/r2r_p2p_tests.log
fi
for h in nccl p2p; do
  if [ -z "$SKIP_TIMELINE" ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
      tools/partition_timeline.py --halo $h > gpurun_out/r2r_timeline_${h}_n$N.txt 2> gpurun_out/r2r_timeline_${h}_n$N.err
  fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
      bench.py --gpus $N --steps 20 --warmup 3 --halo $h --skip-gpu-eager $( [ "$h" = nccl ] && [ "$N" != 2 ] && echo --skip-parity ) > gpurun_out/r2r_mus_${h}_n$N.json 2> gpurun_out/r2r_mus_${h}_n$N.err
done
cat gpurun_out/r2r_p2p_tests.log 2>/dev/null; for f in gpurun_out/r2r_mus_*_n$N.json; do echo $f; head -c 500 $f; echo; done
grep -h "^#" gpurun_out/r2r_timeline_*_n$N.txt; tail -n 3 gpurun_out/r2r_*_n$N.err
