"""profiles/traffic.json from an `ncu --set full` capture of the GEMM-filter launches of ONE bench step.

    python tools/make_traffic.py gpurun_out/gemm_r02k.ncu-rep profiles/r02k_gemm_filter_full.txt > profiles/traffic.json

bench.py reads the file and reports `roofline.traffic` (dram read + write bytes per step, like `achieved`)."""
import csv
import io
import json
import subprocess
import sys

rep, src = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
ri, wi, ti = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
tot_r = sum(float(r[ri]) * mult[units[ri]] for r in rows[2:])
tot_w = sum(float(r[wi]) * mult[units[wi]] for r in rows[2:])
tmul = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[ti]]
print(json.dumps({
    "gemm_filter_dram_bytes_per_step": tot_r + tot_w, "dram_read_bytes": tot_r, "dram_write_bytes": tot_w,
    "launches": len(rows) - 2, "kernel_ms_under_ncu": sum(float(r[ti]) * tmul for r in rows[2:]),
    "algorithmic_shadow_bytes_per_step": 10_000_000 * 208 * 2 * 1.0 + 65536 * 208 * 2,
    "algorithmic_fp32_bytes_per_step": 10_000_000 * 200 * 4,
    "source": f"ncu --set full --clock-control none -k regex:gemm_filter -s 12 -c 6 (one step = 1 seed + 5 chunk launches), {src}",
}, indent=1))
