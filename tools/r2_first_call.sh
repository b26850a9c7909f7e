#!/bin/bash
# First GPU call of the next round (one B200):   gpurun --timeout 1500 -- 'bash tools/r2_first_call.sh'
# 1. the GPU parity suite; 2. the experimental edge-kernel variants (csrc/mp_edge_pair_tma.cu), each in its own process
# under a timeout so that a hang in one mode costs two minutes, not the call; 3. their timings next to the default kernel.
# Everything is written to gpurun_out/r2_*.log.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest.log
# the TMA assumptions in isolation (one-warp kernels): tells which primitive is off if a mode below fails
G4C_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_tma_primitives.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_tma_primitives.log
for m in 1 2 3 4; do
  G4C_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_edge_pair.py -m gpu -x -q -k "tma_modes and mode${m}" > gpurun_out/r2_experimental_mode${m}.full 2>&1
  rc=$?                                   # 124 = the timeout fired (a hang), 1 = a parity failure
  tail -12 gpurun_out/r2_experimental_mode${m}.full > gpurun_out/r2_experimental_mode${m}.log
  echo "mode $m parity exit $rc" >> gpurun_out/r2_experimental_mode${m}.log
done
for m in 1 2 3 4; do
  timeout 120 python tools/bench_edge.py --modes 0,$m 2>&1 | tail -4 > gpurun_out/r2_bench_edge_mode${m}.log
  timeout 120 python tools/bench_edge.py --modes 0,$m --layers 2 --k 6 2>&1 | tail -4 >> gpurun_out/r2_bench_edge_mode${m}.log
done
# memcheck of the small cases of every mode (shared-memory / global out-of-bounds, misaligned bulk copies)
G4C_TEST_EXPERIMENTAL=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_edge_pair.py -m gpu -q \
  -k "tma_modes and 77-6" > gpurun_out/r2_memcheck.full 2>&1
tail -25 gpurun_out/r2_memcheck.full > gpurun_out/r2_memcheck.log
tail -n 20 gpurun_out/r2_*.log
