"""ncu --csv logs (long format: one row per launch x metric) -> a per-kernel table for profiles/ and profiles/ncu_traffic.json.
    python tools/summarize_ncu.py gpurun_out/r2p/proj.csv [--json-key proj_wgrad=proj_wgrad_ts_kernel ...] > profiles/r2_proj_ncu.txt"""
import collections
import csv
import json
import sys

UNIT = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    col = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        d = launches.setdefault(r[col["ID"]], {"name": r[col["Kernel Name"]]})
        try:
            v = float(r[col["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        d[r[col["Metric Name"]]] = v * UNIT.get(r[col["Metric Unit"]], 1.0)
    return list(launches.values())


def main():
    path = sys.argv[1]
    keys = dict(a.split("=") for a in sys.argv[2:] if "=" in a and not a.startswith("--"))
    L = load(path)
    by = collections.OrderedDict()
    for d in L:
        by.setdefault(d["name"].split("(")[0][:70], []).append(d)
    print(f"# {path}: {len(L)} profiled launches; per kernel: mean over launches (ncu replays each kernel: durations are cold-cache, compare SHARES and per-launch bytes)")
    print(f"{'kernel':58s} {'n':>3s} {'us':>9s} {'dram_rd_MB':>10s} {'dram_wr_MB':>10s} {'dram_GB/s':>9s} {'dram%':>6s} {'L2hit%':>6s} {'lts%':>6s} {'tensor%':>7s} {'warps%':>6s} {'regs':>4s} {'xbar2l1_MB':>10s}")
    out_json = {}
    for name, ds in by.items():
        m = lambda k: sum(d.get(k, 0.0) for d in ds) / len(ds)
        us = m("gpu__time_duration.sum")
        rd, wr = m("dram__bytes_read.sum"), m("dram__bytes_write.sum")
        print(f"{name:58s} {len(ds):3d} {us:9.1f} {rd / 1e6:10.1f} {wr / 1e6:10.1f} {(rd + wr) / max(us, 1e-9) / 1e3:9.0f} "
              f"{m('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {m('lts__t_sector_hit_rate.pct'):6.1f} "
              f"{m('lts__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {m('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):7.1f} "
              f"{m('sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} {m('launch__registers_per_thread'):4.0f} {m('l1tex__m_xbar2l1tex_read_bytes.sum') / 1e6:10.1f}")
        for k, pat in keys.items():
            if pat in name:
                out_json[k] = {"bytes": int(rd + wr), "source": f"{path} ({name}, mean of {len(ds)} launches, dram__bytes_read.sum + dram__bytes_write.sum)"}
    if out_json:
        sys.stderr.write(json.dumps(out_json) + "\n")


if __name__ == "__main__":
    main()
