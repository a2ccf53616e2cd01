"""Turns the scratch ncu outputs under gpurun_out/ into the committed summaries under profiles/.

  python tools/summarize_ncu.py <tag> <launches.csv> <mac_report.ncu-rep> [<other_report.ncu-rep> ...]
writes profiles/<tag>_launches.csv (trimmed launch list), profiles/<tag>_<kernel>.json (key counters of each full
capture) and refreshes profiles/mac_kernel_traffic.json (per-launch DRAM bytes of the dominant kernel, read by bench.py).
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def to_unit(v, u, want):
    v = float(v)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}
    return v * scale.get(u, 1)


def launches(tag, path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
        out.append((r[ix["ID"]], name, to_unit(r[ix["Metric Value"]], r[ix["Metric Unit"]], "us")))
    dst = os.path.join(ROOT, "profiles", f"{tag}_launches.csv")
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: the first 400 launches of `python bench.py --steps 2 --warmup 1 --no-cpu --quick`\n")
        f.write("# (cold-cache, serialised: compare SHARES, not absolutes)\nid,kernel,duration_us\n")
        for i, n, v in out:
            f.write(f"{i},{n},{v:.2f}\n")
    # share of the serialised step: medians over all launches of the two kernels of a step (the host-buffer calls at the
    # end of bench.py read w_ccs over PCIe inside the witness kernel, so "the last pair" would not be representative)
    import statistics
    wit = [v for _, n, v in out if "witness_kernel" in n]
    mac = [v for _, n, v in out if "mac_kernel<1" in n.replace(" ", "")]
    if wit and mac:
        w, m = statistics.median(wit), statistics.median(mac)
        print(dst, f"median witness_kernel {w:.1f} us ({100 * w / (w + m):.0f} %), median mac_kernel<1,8> {m:.1f} us ({100 * m / (w + m):.0f} %)")


_written = set()


def report(tag, path):
    txt = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("lat::", "")
        short = re.sub(r"[^a-z0-9_]+", "_", name.lower()).strip("_")
        if short in seen or short in _written:  # first capture of a kernel wins (bench.py before the fold script)
            continue
        seen.add(short)
        _written.add(short)
        out = {"kernel": d["Kernel Name"], "source_report": os.path.basename(path),
               "command": "ncu --set full --clock-control none --import-source on -k regex:<kernels> -c 6 python "
                          + ("tools/run_crt_once.py" if "crt" in os.path.basename(path) else "tools/run_ntt_once.py" if "ntt" in os.path.basename(path) else
                             "tools/run_fold_once.py" if "fold" in os.path.basename(path) else "bench.py --steps 2 --warmup 1 --no-cpu --quick")}
        for k in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                out[k] = {"value": float(d[k]), "unit": u[k]}
        stalls = {}
        for h in hdr:
            if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and d[h] not in ("", "n/a"):
                stalls[h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = round(float(d[h]), 3)
        out["warp_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        rd = to_unit(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"], "byte")
        wr = to_unit(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"], "byte")
        out["dram_bytes_per_launch"] = rd + wr
        dst = os.path.join(ROOT, "profiles", f"{tag}_{short}.json")
        json.dump(out, open(dst, "w"), indent=1)
        print(dst, "dur", out.get("gpu__time_duration.sum"), "dram bytes", rd + wr)
        if "mac_kernel<1" in name.replace(" ", ""):
            json.dump({"kernel": d["Kernel Name"], "dram_bytes_per_launch": rd + wr, "from": f"profiles/{tag}_{short}.json"},
                      open(os.path.join(ROOT, "profiles", "mac_kernel_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag, sys.argv[2])
    for p in sys.argv[3:]:
        report(tag, p)
