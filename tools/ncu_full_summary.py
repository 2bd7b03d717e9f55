"""One-paragraph summary of each kernel in an `ncu --set full` report: duration, pipe utilisation, DRAM traffic, top stall reasons.
   python tools/ncu_full_summary.py report.ncu-rep [...]   (reads them with `ncu -i ... --page raw --csv`)"""
import csv
import io
import re
import subprocess
import sys

KEYS = [("duration us", "gpu__time_duration.sum", "time"), ("SM busy %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("DRAM read MB", "dram__bytes_read.sum", None), ("DRAM write MB", "dram__bytes_write.sum", None),
        ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1), ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
        ("XU %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1), ("ALU %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
        ("FMA %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1), ("LSU %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
        ("regs", "launch__registers_per_thread", 1), ("smem/block KB", "launch__shared_mem_per_block_dynamic", "kb"), ("waves/SM", "launch__waves_per_multiprocessor", 1),
        ("occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1)]
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}

for path in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        print(path, ": unreadable")
        continue
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        name = re.sub(r"\(.*", "", d.get("Kernel Name", ("", "?"))[1]).replace("void ", "")
        print("== %s   grid %s block %s   [%s]" % (name, d.get("Grid Size", ("", ""))[1], d.get("Block Size", ("", ""))[1], path.split("/")[-1]))
        parts = []
        for label, key, scale in KEYS:
            if key in d and d[key][1] not in ("", "n/a"):
                u, v = d[key]
                try:
                    x = float(v.replace(",", ""))
                except ValueError:
                    continue
                if scale == "time":
                    x *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(u, 1.0)
                elif scale == "kb":
                    x *= {"byte/block": 1e-3, "Kbyte/block": 1.0, "Mbyte/block": 1e3}.get(u, 1e-3)
                else:
                    x = x * (UNIT.get(u, 1e-6) if scale is None else scale)
                parts.append("%s %.1f" % (label, x))
        print("   " + " | ".join(parts))
        st = []
        for h, (u, v) in d.items():
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
            if m and "not_issued" not in h:
                try:
                    st.append((float(v.replace(",", "")), m.group(1)))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("   stalls per issue: " + ", ".join("%s %.2f" % (n, x) for x, n in st[:5]))
