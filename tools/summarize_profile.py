"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profile.py launches gpurun_out/launches_r1.csv profiles/r1_decode_launches.md "title"
    python tools/summarize_profile.py rep gpurun_out/prof_gemm_r1.ncu-rep profiles/r1_gemm_full.md "title"
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst, title):
    rows = list(csv.reader(open(src)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[hdr]
    ki, vi, mi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name")
    agg = collections.OrderedDict()
    n = 0
    for r in rows[hdr + 1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("umv::", "")[:80]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
        n += 1
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: "
                f"compare SHARES, not absolutes).  {n} launches, {tot / 1e3:.1f} us total.\n\n")
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {t / 1e3:.1f} | {100 * t / tot:.1f}% | {t / c / 1e3:.2f} |\n")
    print(open(dst).read())


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg"]


def rep(src, dst, title):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    cols = [(w, H.index(w)) for w in WANT if w in H]
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nSource: `ncu --set full --clock-control none --import-source on` ({src.split('/')[-1]}, not tracked); "
                "values per launch.\n\n")
        f.write("| kernel | " + " | ".join(f"{w} [{U[i]}]" for w, i in cols) + " |\n|---|" + "---:|" * len(cols) + "\n")
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[H.index("Kernel Name")]).replace("void ", "").replace("umv::", "")
            f.write(f"| `{name}` | " + " | ".join(r[i] for _, i in cols) + " |\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "rep": rep}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
