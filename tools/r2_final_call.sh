#!/bin/bash
# The driver's round-end sequence on one B200: GPU suite, smoke(), reference arm, our arm.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -rs --durations=6 2>&1 | tail -25 > gpurun_out/${TAG:-r2n}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG:-r2n}_smoke.log 2>&1
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG:-r2n}_bench_ref.json 2> gpurun_out/${TAG:-r2n}_bench_ref.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG:-r2n}_bench.json 2> gpurun_out/${TAG:-r2n}_bench.err
cat gpurun_out/${TAG:-r2n}_pytest.log gpurun_out/${TAG:-r2n}_smoke.log | cut -c1-220; head -c 900 gpurun_out/${TAG:-r2n}_bench.json; echo; head -c 300 gpurun_out/${TAG:-r2n}_bench_ref.json
