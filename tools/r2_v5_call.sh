#!/bin/bash
# GPU call: parity of the v5 edge kernel (csrc/mp_edge_v5.cu), its timing next to v3, and its in-kernel phase profile.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_edge_pair.py tests/test_gpu_tma_primitives.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_edge_tests.log
timeout 120 python tools/bench_edge.py --variants v3,v5 2>&1 | tail -4 > gpurun_out/r2b_bench_edge.log
timeout 120 python tools/bench_edge.py --variants v3,v5 --layers 2 2>&1 | tail -4 >> gpurun_out/r2b_bench_edge.log
timeout 120 python tools/bench_edge.py --variants v5 --no-e 2>&1 | tail -4 >> gpurun_out/r2b_bench_edge.log
G4C_PROFILE=1 G4C_LIB=$PWD/graphs4cfd_b200/libg4c_prof.so timeout 120 python tools/bench_edge.py --variants v5 2>&1 | tail -6 > gpurun_out/r2b_phases_v5.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2b_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -n 30 gpurun_out/r2b_*.log gpurun_out/r2b_bench.json
