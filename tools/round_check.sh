#!/bin/bash
# Verification + evidence pass at HEAD on one B200 (run under gpurun; outputs land in gpurun_out/ with tag $1):
#   GPU test suite, smoke, per-chunk debug timing, the default bench line, the reference arm, the ncu launch list of
#   the bench command and one `ncu --set full` capture of the dominant kernel's launches of one step.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --durations=12 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo PYTEST_EXIT $?; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
B200ANN_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_debug_$TAG.log 2>&1; echo DEBUG_EXIT $?
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.log 2>&1; echo BENCH_EXIT $?
tail -1 gpurun_out/bench_$TAG.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; echo REF_EXIT $?; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-300
# the ncu passes profile the headline step only (--no-extra: no sub-runs); every number printed under ncu is ignored
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_launches_$TAG.log 2>&1; echo NCU_LAUNCHES_EXIT $?
# SKIP_FULL=1: no `--set full` captures (for a pass whose kernels are byte-identical to the previous capture's)
if [ -n "$SKIP_FULL" ]; then ls -la gpurun_out | tail -12; exit 0; fi
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_filter -s 12 -c 6 -f -o gpurun_out/gemm_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_full_$TAG.log 2>&1; echo NCU_FULL_EXIT $?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalize_kernel -s 3 -c 1 -f -o gpurun_out/finalize_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_finalize_$TAG.log 2>&1; echo NCU_FINALIZE_EXIT $?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 4 -c 1 -f -o gpurun_out/scan_$TAG \
    python tools/scan_profile_target.py > gpurun_out/ncu_scan_$TAG.log 2>&1; echo NCU_SCAN_EXIT $?
ls -la gpurun_out | tail -12
