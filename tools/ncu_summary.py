"""Summarise ncu outputs brought back in gpurun_out/ into text files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r01b.csv  > profiles/r01b_launches_summary.txt
    python tools/ncu_summary.py full gpurun_out/gemm_r01b.ncu-rep      > profiles/r01b_gemm_filter_full.txt
"""
import collections
import csv
import io
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v * 1e6 if r[ui] == "s" else v
        a = agg.setdefault(r[ki][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(data)} launches, {tot / 1e3:.2f} ms total under ncu (cold-cache, serialised: compare SHARES)")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t / 1e3:10.3f} ms  {c:4d}x  {100 * t / tot:5.1f}%  {n}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    ki = hdr.index("Kernel Name")
    print(f"# {path}: ncu --set full, per launch")
    for r in rows[2:]:
        print(f"\n== {r[ki][:100]}")
        for w, i in idx:
            print(f"   {w:95s} {r[i]:>14s} {units[i]}")
        try:
            rd = float(r[hdr.index('dram__bytes_read.sum')]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index('dram__bytes_read.sum')]]
            wr = float(r[hdr.index('dram__bytes_write.sum')]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index('dram__bytes_write.sum')]]
            print(f"   {'traffic = dram read + write (bytes)':95s} {rd + wr:14.0f}")
        except Exception:
            pass


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
