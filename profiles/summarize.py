"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

  python profiles/summarize.py launches gpurun_out/launches_r01.csv  > profiles/launches_r01.txt
  python profiles/summarize.py raw      gpurun_out/top_r01.ncu-rep   > profiles/top_r01.txt
  python profiles/summarize.py traffic  gpurun_out/frame_r01b.ncu-rep        (JSON: DRAM bytes per launch and kernel)
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        agg[re.sub(r"\(.*", "", row["Kernel Name"])[:80]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({path})")
    print("# (serialised, cold-cache launches: compare SHARES, not absolutes)")
    print(f"# total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")
    print(f"{'kernel':82s} {'n':>5s} {'mean_us':>9s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:82s} {len(v):5d} {sum(v) / len(v):9.2f} {100 * sum(v) / tot:6.1f}%")


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# selected metrics from `ncu --set full --clock-control none` ({path}); one block per captured launch")
    for row in rows[2:]:
        print("kernel:", re.sub(r"\(.*", "", row[hdr.index("Kernel Name")])[:90])
        for w in WANT:
            if w in hdr:
                print(f"    {w:62s} {row[hdr.index(w)]:>16s} {units[hdr.index(w)]}")


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def traffic(path):
    """dram__bytes_read.sum + dram__bytes_write.sum per captured launch (ncu prints per-row units in the second row only,
    so this re-queries with --print-units base)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        g = lambda k: float(row[hdr.index(k)].replace(",", "")) * UNIT.get(units[hdr.index(k)], 1.0)
        res.append({"kernel": re.sub(r"\(.*", "", row[hdr.index("Kernel Name")]), "grid": int(float(row[hdr.index("launch__grid_size")])),
                    "duration_us": g("gpu__time_duration.sum") / (1e3 if units[hdr.index("gpu__time_duration.sum")] in ("ns", "nsecond") else 1.0),
                    "dram_read_bytes": g("dram__bytes_read.sum"), "dram_write_bytes": g("dram__bytes_write.sum")})
    print(json.dumps(res, indent=1))


def stalls(path, top=16):
    """Warp-state samples per CUDA source line (needs --import-source on and -lineinfo): the lines where the warps wait."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    fname, hdr, rows = "?", None, []
    for r in csv.reader(out.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif len(r) > 10 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].strip().isdigit():
            g = lambda k: int(float(r[hdr.index(k)] or 0)) if r[hdr.index(k)] not in ("-", "") else 0
            rows.append((g("# Samples"), g("stall_barrier"), g("stall_long_sb"), g("stall_wait"), g("stall_short_sb"), g("stall_branch_resolving"),
                         g("Instructions Executed"), f"{fname}:{r[0]}", r[1].strip()[:56]))
    merged = {}                                   # several captured launches of the kernel: one entry per source line
    for x in rows:
        m = merged.setdefault(x[7], [0, 0, 0, 0, 0, 0, 0, x[7], x[8]])
        for i in range(7):
            m[i] += x[i]
    rows = [tuple(v) for v in merged.values()]
    tot = sum(x[0] for x in rows) or 1
    print(f"total samples {tot} barrier {sum(x[1] for x in rows)} long_sb {sum(x[2] for x in rows)} wait {sum(x[3] for x in rows)} "
          f"short {sum(x[4] for x in rows)} branch {sum(x[5] for x in rows)}")
    for x in sorted(rows, reverse=True)[:top]:
        print(f"{100 * x[0] / tot:5.1f}% bar={x[1]:6d} lsb={x[2]:6d} wait={x[3]:6d} ssb={x[4]:5d} br={x[5]:5d} inst={x[6]:9d}  {x[7]}  {x[8]}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "traffic": traffic, "stalls": stalls}[sys.argv[1]](sys.argv[2])
