"""Turn gpurun_out/<round>_* ncu outputs into the tracked summaries under profiles/."""
import collections, csv, json, os, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(pr, exist_ok=True)
out = ["# %s profile summary" % R, "",
       "Source: `tools/make_profiles.sh %s` under gpurun (1x B200); raw launch list kept as `%s_bench_launches.csv`." % (R, R), ""]
# ---- launch list
f = os.path.join(go, "%s_bench_launches.csv" % R)
if os.path.exists(f):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    h = rows[0]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")) / {"ns": 1e3, "us": 1.0, "ms": 1e-3}.get(r[ui], 1e3)
        k = r[ki].split("(")[0][:70]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out += ["## Launch list of `python bench.py --steps 1 --warmup 3 --batch 4` (ncu gpu__time_duration, cold cache, serialised)", "",
            "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        out.append("| `%s` | %d | %.0f | %.1f %% |" % (k, a[0], a[1], 100 * a[1] / tot))
    out += ["", "total %d launches, %.1f ms of kernel time" % (sum(a[0] for a in agg.values()), tot / 1e3), ""]
    import shutil; shutil.copy(f, os.path.join(pr, os.path.basename(f)))
# ---- full capture
f = os.path.join(go, "%s_hot_raw.csv" % R)
if os.path.exists(f):
    import shutil; shutil.copy(f, os.path.join(pr, os.path.basename(f)))  # raw capture travels with its summary
    rows = list(csv.reader(open(f)))
    h = rows[0]
    want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
            ("sm__inst_executed_pipe_tensor.sum", "tensor inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("launch__registers_per_thread", "regs"),
            ("launch__grid_size", "grid"), ("launch__block_size", "block")]
    cols = [(h.index(k), n) for k, n in want if k in h]
    out += ["## `ncu --set full` of the hot kernels (B=4, 15x320x320; units row: %s)" % ", ".join(
        "%s [%s]" % (n, rows[1][i]) for i, n in cols), "", "| kernel | " + " | ".join(n for _, n in cols) + " |",
        "|---|" + "---:|" * len(cols)]
    ki = h.index("Kernel Name")
    for r in rows[2:]:
        out.append("| `%s` | " % r[ki].split("(")[0][:40] + " | ".join(r[i][:12] for i, _ in cols) + " |")
    out.append("")
    # DRAM traffic of the DC gradient per launch (bench.py's roofline.traffic reads this file)
    try:
        ri, wi = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        def dram(r):
            return float(r[ri].replace(",", "")) * mult[rows[1][ri]] + float(r[wi].replace(",", "")) * mult[rows[1][wi]]
        dc = [r for r in rows[2:] if "row_dc" in r[ki]]
        # the regulariser kernels of ONE time step = the launches between two DC gradients of the capture
        idx = [i for i, r in enumerate(rows[2:]) if "row_dc" in r[ki]]
        stack = rows[2:][idx[0] + 1:idx[1]] if len(idx) >= 2 else [r for r in rows[2:] if "row_dc" not in r[ki]]
        tj = {}
        if dc:
            tot = sum(dram(r) for r in dc) / len(dc)
            tj["row_dc"] = {"dram_bytes_per_launch": tot, "slices": 4, "kernel": dc[0][ki].split("(")[0],
                            "source": "%s_hot_raw.csv (ncu --set full, B=4, 15x320x320)" % R}
            out += ["DC gradient DRAM traffic per launch (B=4): %.1f MB (algorithmic: 108.1 MB)" % (tot / 1e6), ""]
        if stack:
            tots = sum(dram(r) for r in stack)
            tj["conv_stack"] = {"dram_bytes_per_time_step": tots, "slices": 4, "launches": len(stack),
                                "kernels": [r[ki].split("(")[0][:40] for r in stack],
                                "source": "%s_hot_raw.csv (ncu --set full, B=4, 15x320x320)" % R}
            out += ["Regulariser DRAM traffic per time step (B=4, %d launches): %.1f MB" % (len(stack), tots / 1e6), ""]
        if tj:
            json.dump(tj, open(os.path.join(pr, "%s_traffic.json" % R), "w"))
    except ValueError:
        pass
f = os.path.join(go, "%s_bench.json" % R)
if os.path.exists(f):
    line = open(f).read().strip().splitlines()[-1]
    out += ["## bench line measured in the same gpurun call (outside the profiler)", "", "```json", line, "```", ""]
open(os.path.join(pr, "%s_summary.md" % R), "w").write("\n".join(out))
print("\n".join(out)[:6000])
