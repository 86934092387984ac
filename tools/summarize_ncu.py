"""Turn an .ncu-rep (ncu --set full) or a launch-list CSV (ncu --metrics gpu__time_duration.sum --csv)
into the small text summaries committed under profiles/.

    python tools/summarize_ncu.py rep   gpurun_out/prof.ncu-rep      > profiles/r01_xxx_full.txt
    python tools/summarize_ncu.py list  gpurun_out/launches.csv      > profiles/r01_xxx_launches.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    idx = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    print("# source: %s  (ncu --set full --clock-control none; one block per profiled launch)" % path)
    for r in rows[2:]:
        print("\n== %s" % r[ki])
        for m, i in idx:
            print("  %-72s %-16s %s" % (m, units[i], r[i]))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]); ru = units[hdr.index("dram__bytes_read.sum")]
            wr = float(r[hdr.index("dram__bytes_write.sum")]); wu = units[hdr.index("dram__bytes_write.sum")]
            sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            print("  %-72s %-16s %.1f" % ("dram traffic (read + write)", "Mbyte", (rd * sc[ru] + wr * sc[wu]) / 1e6))
        except Exception:
            pass


def launch_list(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
        k = re.sub(r"\(.*", "", r["Kernel Name"]).strip() + " grid=" + r["Grid Size"].replace(" ", "")
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    print("# source: %s  (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised:" % path)
    print("# compare SHARES, not absolutes)   launches=%d   total=%.1f us" % (len(rows), T))
    print("%-78s %6s %12s %10s %7s" % ("kernel", "n", "sum us", "avg us", "share"))
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print("%-78s %6d %12.1f %10.2f %6.1f%%" % (k[:78], cnt[k], v, v / cnt[k], 100 * v / T))
    fam = collections.defaultdict(float)
    for k, v in tot.items():
        fam[k.split("<")[0].split(" grid")[0]] += v
    print("\nby kernel family:")
    for k, v in sorted(fam.items(), key=lambda x: -x[1]):
        print("  %-60s %6.1f%%" % (k[:60], 100 * v / T))


if __name__ == "__main__":
    {"rep": rep, "list": launch_list}[sys.argv[1]](sys.argv[2])
