"""Turn ncu outputs into the small, tracked summaries under profiles/.
  python scripts/summarize_ncu.py full  gpurun_out/prof.ncu-rep  profiles/r01_pull_tile_full.md  [title]
  python scripts/summarize_ncu.py list  gpurun_out/launches.csv  profiles/r01_launches.md        [title]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum",
    "l1tex__m_xbar2l1tex_read_sectors_mem_global_op_tma_ld.sum",
    "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
]


def full(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# %s" % title, "", "Source: `ncu --set full --clock-control none --import-source on` (report `%s`, not tracked)." % rep, ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines += ["## %s" % name, "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append("| %s | %s | %s |" % (k, r[i], units[i]))
        rd = float(r[hdr.index("dram__bytes_read.sum")]) if "dram__bytes_read.sum" in hdr else 0
        wr = float(r[hdr.index("dram__bytes_write.sum")]) if "dram__bytes_write.sum" in hdr else 0
        u1 = units[hdr.index("dram__bytes_read.sum")]
        u2 = units[hdr.index("dram__bytes_write.sum")]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        tot = rd * scale.get(u1, 1) + wr * scale.get(u2, 1)
        lines += ["", "DRAM traffic per launch (read + write): **%.3f GB**" % (tot / 1e9), "", "Warp stall samples (pc sampling):", ""]
        st = []
        for i, k in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
                try:
                    st.append((int(float(r[i])), k.split("stalled_")[1]))
                except ValueError:
                    pass
        tot_s = sum(v for v, _ in st) or 1
        for v, k in sorted(st, reverse=True)[:8]:
            lines.append("* %s: %.1f %%" % (k, 100.0 * v / tot_s))
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")


def launch_list(path, out, title):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        unit = r[hdr.index("Metric Unit")]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values()) or 1
    lines = ["# %s" % title, "", "Source: `ncu --metrics gpu__time_duration.sum --clock-control none` over `%s` "
             "(cold-cache, serialised launches: compare SHARES, not absolutes)." % path, "",
             "| kernel | launches | total ms | share | avg ms |", "|---|---|---|---|---|"]
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.3f | %.1f %% | %.4f |" % (name[:90], n, ns / 1e6, 100 * ns / tot, ns / 1e6 / n))
    # the kernels launched every iteration, as shares of an iteration (graph construction excluded)
    it = collections.OrderedDict((k, v) for k, v in agg.items() if any(t in k for t in (
        "seg_tile_kernel", "pull_tile_kernel", "pull_fixup", "combine_hub", "hot_refresh", "hot_permute", "empties_kernel", "pack_values")))
    tot_it = sum(a[1] for a in it.values()) or 1
    lines += ["", "Per-iteration kernels only (shares of one iteration's device time):", "",
              "| kernel | launches | total ms | share of iteration | avg ms |", "|---|---|---|---|---|"]
    for name, (n, ns) in sorted(it.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.3f | %.1f %% | %.4f |" % (name[:110], n, ns / 1e6, 100 * ns / tot_it, ns / 1e6 / n))
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else dst
    (full if mode == "full" else launch_list)(src, dst, title)
