#!/bin/bash
# One-GPU validation call: the whole GPU suite (incl. the drop-in, long-rollout and renumbering tests), the bench line of both
# arms (the reference arm carries gpu_eager), the REMuS bench line.  Everything lands in gpurun_out/r2f_*.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -rs --durations=8 2>&1 | tail -40 > gpurun_out/r2f_pytest.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --weights init --skip-cpu-baseline > gpurun_out/r2f_bench_init_weights.json 2> gpurun_out/r2f_bench_init_weights.err
timeout 900 python bench.py --model remus --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/r2f_bench_remus.json 2> gpurun_out/r2f_bench_remus.err
tail -n 40 gpurun_out/r2f_pytest.log; tail -c 1500 gpurun_out/r2f_bench_ref.json; tail -c 600 gpurun_out/r2f_bench.err; tail -c 400 gpurun_out/r2f_bench_remus.err
