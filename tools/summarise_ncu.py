"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarise_ncu.py launches <launches.csv> <out.md> [title]
  python tools/summarise_ncu.py full <prof.ncu-rep> <out.md> [--traffic]

`launches`: per-kernel launch counts, total / mean device time and SHARE of the run.
`full`: for every distinct kernel of the report (its longest launch): key `ncu --set full` metrics, warp
stall reasons, and -- for single-kernel reports captured with --import-source -- the SASS opcode mix
weighted by executed count.  --traffic also rewrites profiles/integrate_kernel_traffic.json (DRAM bytes
per launch, read by bench.py for roofline.traffic).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

NAMES = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
         "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
         "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum",
         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
         "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
         "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def short(name):
    return name.split("(")[0].replace("otslam::", "").replace("void ", "")


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        agg[short(r[ki])].append(float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1e-3))
    tot = sum(sum(v) for v in agg.values())
    lines = [f"# ncu launch list: {title}\n", "`ncu --metrics gpu__time_duration.sum --clock-control none`, our kernels only. "
             "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n",
             "| kernel | launches | total us | mean us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{k}` | {len(v)} | {sum(v):.1f} | {sum(v)/len(v):.1f} | {sum(v)/tot:.3f} |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def full(path, out, traffic):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki, ti = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
    best = collections.OrderedDict()
    for r in rows[2:]:
        k = short(r[ki])
        if k not in best or float(r[ti].replace(",", "")) > float(best[k][ti].replace(",", "")):
            best[k] = r
    lines = [f"# ncu --set full --clock-control none: {os.path.basename(path)}\n",
             "One section per distinct kernel of the capture (its longest launch).\n"]
    for k, R in best.items():
        lines += [f"## `{k}`\n", "| metric | value | unit |", "|---|---:|---|"]
        for n in NAMES:
            if n in hdr:
                lines.append(f"| `{n}` | {R[hdr.index(n)]} | {units[hdr.index(n)]} |")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(R[i].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        lines += ["\nwarp stall reasons (warps per issue-active cycle): " +
                  ", ".join(f"{h} {v:.2f}" for v, h in sorted(stalls, reverse=True)[:7]) + "\n"]
    if len(best) == 1:
        src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        hi = [i for i, r in enumerate(srows) if "Source" in r and "Instructions Executed" in r]
        if hi:
            h2 = srows[hi[-1]]
            si, ei = h2.index("Source"), h2.index("Instructions Executed")
            ops, tot = collections.Counter(), 0
            for r in srows[hi[-1] + 1:]:
                try:
                    n = int(r[ei])
                except (ValueError, IndexError):
                    continue
                t = r[si].split()
                if not t:
                    continue
                op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
                ops[op] += n
                tot += n
            if tot:
                lines += ["## SASS opcode mix (warp instructions executed)\n", "| opcode | executed | share |", "|---|---:|---:|"]
                lines += [f"| {k} | {v} | {100*v/tot:.1f}% |" for k, v in ops.most_common(16)]
                tma = [k for k in ops if k in ("UBLKCP", "SYNCS", "UTMALDG", "UTMASTG", "UTMAPF")]
                lines.append(f"\nTMA / mbarrier opcodes present: {', '.join(sorted(tma)) or 'none'}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if traffic:
        k, R = next(iter(best.items()))
        val = lambda n: float(R[hdr.index(n)].replace(",", "")) * SCALE.get(units[hdr.index(n)], 1)  # noqa: E731
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        json.dump({"kernel": k, "source": os.path.basename(path), "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "dram_bytes_per_launch": rd + wr, "launch_us": val("gpu__time_duration.sum"),
                   "issue_active_frac": float(R[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")].replace(",", "")) / 100.0,
                   "grid_size": int(float(R[hdr.index("launch__grid_size")].replace(",", ""))),
                   "registers_per_thread": int(float(R[hdr.index("launch__registers_per_thread")].replace(",", ""))),
                   "note": "one launch = one 32-frame batch of the default bench workload (configs[1]: 640x480 / 5 mm chair+table)"},
                  open(os.path.join(ROOT, "profiles", "integrate_kernel_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else os.path.basename(sys.argv[2]))
    else:
        full(sys.argv[2], sys.argv[3], "--traffic" in sys.argv)
