#!/usr/bin/env python
"""Condenses `ncu --page raw --csv` exports (gpurun_out/r2_ncu_*_raw.csv) into the evidence kept under profiles/:
one row per captured launch with the columns the roofline needs, and a per-kernel JSON of DRAM bytes per launch that
bench.py reads for `roofline.traffic`.

    python tools/ncu_summary.py gpurun_out/r2_ncu_conv_v27_raw.csv ... --tag r02_v27
"""
import csv
import json
import re
import sys
from collections import OrderedDict

COLS = OrderedDict([
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_ncu_peak"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_MB"),
    ("lts__t_bytes.sum", "l2_bytes_MB"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1_lsu_wavefront_pct"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_tensor_read_wavefront_pct"),
    ("sm__inst_executed_pipe_tc.sum.pct_of_peak_sustained_active", "pipe_tc_inst_pct"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "pipe_tensor_cycles_active_pct"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "mem_tensor_cycles_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
])


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("ofb::", "")


def main():
    tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else "r02"
    args = [a for a in sys.argv[1:] if not a.startswith("--") and a != tag]
    rows_out, per_kernel = [], OrderedDict()
    for path in args:
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ki = hdr.index("Kernel Name")
        idx = [(hdr.index(c) if c in hdr else None) for c in COLS]
        for r in data:
            vals = []
            for (c, _), i in zip(COLS.items(), idx):
                v = r[i] if i is not None and r[i] != "" else ""
                if v and units[i] in ("byte", "Kbyte", "Gbyte"):
                    v = str(float(v) * {"byte": 1e-6, "Kbyte": 1e-3, "Gbyte": 1e3}[units[i]])
                if v and c == "gpu__time_duration.sum" and units[i] in ("ns", "ms"):
                    v = str(float(v) * {"ns": 1e-3, "ms": 1e3}[units[i]])
                vals.append(v)
            k = short(r[ki])
            rows_out.append([k] + vals)
            d = per_kernel.setdefault(k, {"launches_captured": 0, "time_us": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0, "l2_to_sm_MB": 0.0})
            d["launches_captured"] += 1
            for key, col in (("time_us", 0), ("dram_read_MB", 1), ("dram_write_MB", 2), ("l2_to_sm_MB", 4)):
                d[key] += float(vals[col] or 0)
    with open(f"profiles/{tag}_ncu_launches.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + list(COLS.values()))
        for r in rows_out:
            w.writerow([r[0]] + [(f"{float(v):.4g}" if v else "") for v in r[1:]])
    out = {"source": "ncu --set full --clock-control none of tools/one_forward.py 8 1 (B=8, 512x1024, nrows=4, 2-iter, confidence); "
                     f"per-launch rows in profiles/{tag}_ncu_launches.csv", "kernels": OrderedDict()}
    for k, d in per_kernel.items():
        n = d["launches_captured"]
        out["kernels"][k] = {"launches_captured": n, "avg_us_under_ncu": d["time_us"] / n,
                             "dram_read_mb_per_launch": d["dram_read_MB"] / n, "dram_write_mb_per_launch": d["dram_write_MB"] / n,
                             "l2_to_sm_mb_per_launch": d["l2_to_sm_MB"] / n}
    json.dump(out, open(f"profiles/{tag}_ncu_traffic.json", "w"), indent=1)
    for k, v in out["kernels"].items():
        print(f"{k[:70]:70s} n={v['launches_captured']:3d} {v['avg_us_under_ncu']:8.1f} us  dram {v['dram_read_mb_per_launch']:8.1f} + {v['dram_write_mb_per_launch']:8.1f} MB  l2->sm {v['l2_to_sm_mb_per_launch']:8.1f} MB")


if __name__ == "__main__":
    main()
