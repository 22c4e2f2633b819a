"""A/B of an environment knob on bench.py at several corpus sizes (developer tool; run under gpurun).

    python tools/ab_env.py B200ANN_QSTAGES3 1250000 10000000
"""
import json
import os
import subprocess
import sys

knob, sizes = sys.argv[1], [int(x) for x in sys.argv[2:]] or [10_000_000]
for rows in sizes:
    for on in (0, 1, 0, 1):
        env = dict(os.environ)
        env.pop(knob, None)
        if on:
            env[knob] = "1"
        out = subprocess.run([sys.executable, "bench.py", "--rows", str(rows), "--steps", "20", "--no-extra", "--no-cpu-baseline"],
                             capture_output=True, text=True, env=env).stdout.strip().splitlines()
        d = json.loads(out[-1])
        print(f"rows={rows} {knob}={on}: {d['value']:.0f} q/s {d['ms_per_step']:.3f} ms filter {d['roofline']['achieved']:.0f} TFLOP/s "
              f"digest {d['result_digest'][:8]} clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']} "
              f"{ {k: round(v, 3) for k, v in d['breakdown'].items()} }", flush=True)
