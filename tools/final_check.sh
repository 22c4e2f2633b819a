#!/bin/bash
# End-of-round verification on a B200 (run under gpurun): GPU test suite, smoke, default bench, reference arm.
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo PYTEST_EXIT $?; tail -2 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1; echo BENCH_EXIT $?
tail -1 gpurun_out/bench_final.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo REF_EXIT $?; tail -1 gpurun_out/bench_ref.log | cut -c1-220
