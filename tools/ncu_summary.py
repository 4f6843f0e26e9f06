"""Summaries of an ncu report for profiles/:
   ncu_summary.py launches <launches.csv>                  -> kernel | launches | total ms | share  (from the
                                                             `--metrics gpu__time_duration.sum` launch list)
   ncu_summary.py full <report.ncu-rep> [traffic.json cfg spp_per_wave]
                                                           -> one row per launch of the --set full capture, and
                                                             optionally the dram bytes per k_trace<closest> launch
                                                             that bench.py reports as roofline.traffic"""
import collections
import csv
import io
import json
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Value" in r)
    hdr = rows[hi]
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) != len(hdr):
            continue
        ms = float(r[v].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[u], 1e-6)
        name = r[k].split("(")[0]
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in tot.values())
    print("kernel | launches | total ms | share")
    for name, a in sorted(tot.items(), key=lambda x: -x[1][1]):
        print(f"{name} | {a[0]} | {a[1]:.3f} | {a[1] / total * 100:.1f}%")


COLS = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "IPC"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("l1tex__t_sector_hit_rate.pct", "L1hit%"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_lsu_wavefronts%")]


def full(rep, traffic=None, cfg=None, spp_per_wave=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print("kernel | " + " | ".join(f"{n}[{units[ix[m]]}]" for m, n in COLS if m in ix))
    per = []
    for r in data:
        name = r[ix["Kernel Name"]][:26]
        print(name + " | " + " | ".join(r[ix[m]] for m, _ in COLS if m in ix))
        kn = r[ix["Kernel Name"]]
        if any(s in kn for s in ("k_trace<0>", "k_trace<(bool)0>", "k_trace<false>", "k_trace<1, 0, 0>", "k_trace<(bool)1, (bool)0, (bool)0>")):
            def to_bytes(m):
                return float(r[ix[m]].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[ix[m]]]
            ms = float(r[ix["gpu__time_duration.sum"]]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[units[ix["gpu__time_duration.sum"]]]
            per.append({"dram_bytes": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"), "ms": ms})
    if traffic and per:
        j = {"source": rep, "config": cfg, "spp_per_wave": int(spp_per_wave),
             "k_trace_closest": {"launches": len(per), "dram_bytes_per_launch_avg": sum(p["dram_bytes"] for p in per) / len(per),
                                 "per_launch": per}}
        with open(traffic, "w") as f:
            json.dump(j, f, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(*sys.argv[2:])
