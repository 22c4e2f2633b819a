"""A/B of two builds of the library on bench.py (developer tool; run under gpurun):

    python tools/ab_lib.py the-algorithm_b200/lib/libb200ann_base.so[,other.so ...] [rows ...]

Alternates the in-tree library with the given ones (B200ANN_LIB) and prints value, filter TFLOP/s and the breakdown."""
import json
import os
import subprocess
import sys

others, sizes = sys.argv[1].split(","), [int(x) for x in sys.argv[2:]] or [10_000_000]
for rows in sizes:
    for which in (["new"] + others) * 2:
        env = dict(os.environ)
        env.pop("B200ANN_LIB", None)
        if which != "new":
            env["B200ANN_LIB"] = os.path.abspath(which)
            which = os.path.basename(which).replace("libb200ann_", "").replace(".so", "")
        out = subprocess.run([sys.executable, "bench.py", "--rows", str(rows), "--steps", "20", "--no-extra", "--no-cpu-baseline"],
                             capture_output=True, text=True, env=env).stdout.strip().splitlines()
        d = json.loads(out[-1])
        print(f"rows={rows} lib={which}: {d['value']:.0f} q/s {d['ms_per_step']:.3f} ms filter {d['roofline']['achieved']:.0f} TFLOP/s "
              f"digest {d['result_digest'][:8]} clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']} "
              f"{ {k: round(v, 3) for k, v in d['breakdown'].items()} }", flush=True)
