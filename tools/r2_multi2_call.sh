#!/bin/bash
# Two-GPU call: NCCL parity tests (MuS + REMuS partitions, device-following blocks), then bench lines at N = 2 with the
# in-bench parity check (gathered prediction vs the single-GPU fp32 engine).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_zz_remus_partition.py -m gpu -q -rs -s 2>&1 | grep -E "world=|passed|failed|skipped|Error|error" | tail -30 > gpurun_out/r2i_multi2_pytest.log
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 \
      bench.py --gpus 2 --steps 10 --warmup 3 "$@" > gpurun_out/r2i_${name}_n2.json 2> gpurun_out/r2i_${name}_n2.err; }
run mus
run mus_overlap --overlap --skip-parity
run remus --model remus
cat gpurun_out/r2i_multi2_pytest.log; for f in gpurun_out/r2i_*_n2.json; do echo $f; head -c 700 $f; echo; done; tail -5 gpurun_out/r2i_*_n2.err
