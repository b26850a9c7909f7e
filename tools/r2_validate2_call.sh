#!/bin/bash
# One-GPU call: the whole GPU suite again (after the accelerate() fix), then ncu: the launch list of one bench step and one
# `--set full` capture each of the level-1 MuS edge launch (3 layers) and of the REMuS 36M-angle launch (2 layers).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs --durations=8 2>&1 | tail -60 > gpurun_out/r2g_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu-baseline > gpurun_out/r2g_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_v5 -s 3 -c 1 -o gpurun_out/r2g_edge_v5 \
    python tools/bench_edge.py --variants v5 --reps 1 > gpurun_out/r2g_ncu_v5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:edge_v5 -s 3 -c 1 -o gpurun_out/r2g_edge_v5_remus \
    python tools/bench_edge.py --variants v5 --reps 1 --nodes 6000000 --layers 2 > gpurun_out/r2g_ncu_v5_remus.log 2>&1
tail -n 60 gpurun_out/r2g_pytest.log; tail -3 gpurun_out/r2g_ncu_v5.log gpurun_out/r2g_ncu_v5_remus.log; ls -la gpurun_out/*.ncu-rep | tail -3
