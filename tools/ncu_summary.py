"""Turns ncu output brought back from the GPU box into the small text summaries kept under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches.csv   # per-kernel share of device time
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep       # key metrics of a --set full capture"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "smsp__inst_executed.sum",
]


def launches(path):
    agg = defaultdict(lambda: [0.0, 0])
    with open(path) as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    for r in rows[1:]:
        name = r[kn].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        agg[name][0] += float(r[mv].replace(",", "")) / 1e3
        agg[name][1] += 1
    tot = sum(v[0] for v in agg.values())
    print("# %s: %d launches, %.1f us total (ncu gpu__time_duration.sum, cold-cache & serialised: compare SHARES)"
          % (path, sum(v[1] for v in agg.values()), tot))
    print("%-44s %6s %12s %7s" % ("kernel", "n", "total us", "share"))
    for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%-44s %6d %12.1f %6.1f%%" % (name[:44], n, us, 100 * us / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== %s  grid=%s block=%s" % (d.get("Kernel Name", "?")[:90], d.get("Grid Size"), d.get("Block Size")))
        for k in KEYS:
            if k in d:
                print("   %-66s %14s %s" % (k, d[k], units[hdr.index(k)]))


def traffic(path, kernel_regex, algorithmic_bytes, out="profiles/roofline_traffic.json"):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first launch matching kernel_regex -> the JSON bench.py reads for
    roofline.traffic.  python tools/ncu_summary.py traffic rep.ncu-rep 'dfu_tc_gemm_kernel' 203.0e6"""
    import json
    import re
    out_csv = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out_csv.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if not re.search(kernel_regex, d.get("Kernel Name", "")):
            continue

        def gbytes(key):
            v = float(d[key].replace(",", ""))
            u = units[hdr.index(key)].lower()
            return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        total = gbytes("dram__bytes_read.sum") + gbytes("dram__bytes_write.sum")
        rec = {"bytes_per_launch": total, "algorithmic_bytes_per_launch": float(algorithmic_bytes), "kernel": d["Kernel Name"][:120],
               "grid": d.get("Grid Size"), "source": path, "how": "ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum, one launch"}
        with open(out, "w") as f:
            json.dump(rec, f, indent=1)
        print(json.dumps(rec))
        return
    raise SystemExit("no kernel matching %r in %s" % (kernel_regex, path))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
