"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarise_ncu.py <tag> [launches.csv] [prof.ncu-rep]

writes profiles/launches_<tag>.md (per-kernel launch counts, total / mean device time and SHARE of
the step), profiles/integrate_<tag>.md (key `ncu --set full` metrics + stall reasons + SASS opcode
mix weighted by executed count) and profiles/integrate_kernel_traffic.json (DRAM bytes per launch,
read by bench.py for roofline.traffic).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path, tag):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        agg[r[ki].split("(")[0].replace("otslam::", "")].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = [f"# ncu launch list ({tag}): `--metrics gpu__time_duration.sum --clock-control none`, our kernels only\n",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n",
           "| kernel | launches | total us | mean us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k}` | {len(v)} | {sum(v)/1e3:.1f} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/tot:.3f} |")
    open(os.path.join(ROOT, "profiles", f"launches_{tag}.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def full(path, tag):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, R = rows[0], rows[-1]
    g = lambda n: R[hdr.index(n)] if n in hdr else None  # noqa: E731
    names = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
             "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
             "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
             "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
             "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]
    units = rows[1] if len(rows) > 2 else [""] * len(hdr)
    out = [f"# ncu --set full: integrate_kernel ({tag})\n", "| metric | value | unit |", "|---|---:|---|"]
    for n in names:
        if n in hdr:
            out.append(f"| `{n}` | {g(n)} | {units[hdr.index(n)]} |")
    stalls = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                stalls.append((float(R[i].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    out += ["\n## warp stall reasons (warps per issue-active cycle)\n", "| reason | ratio |", "|---|---:|"]
    out += [f"| {h} | {v:.3f} |" for v, h in sorted(stalls, reverse=True)[:9]]
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
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] += n
            tot += n
        out += ["\n## SASS opcode mix (warp instructions executed)\n", "| opcode | executed | share |", "|---|---:|---:|"]
        out += [f"| {k} | {v} | {100*v/tot:.1f}% |" for k, v in ops.most_common(16)]
        tma = [k for k in ops if k in ("UBLKCP", "SYNCS", "UTMALDG", "UTMASTG")]
        out.append(f"\nTMA / mbarrier opcodes present: {', '.join(sorted(tma)) or 'none'}")
    open(os.path.join(ROOT, "profiles", f"integrate_{tag}.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
    ui = units[hdr.index("dram__bytes_read.sum")]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(ui, 1)
    json.dump({"kernel": "integrate_kernel", "source": os.path.basename(path), "tag": tag,
               "dram_bytes_read": float(rd) * scale, "dram_bytes_write": float(wr) * scale,
               "dram_bytes_per_launch": (float(rd) + float(wr)) * scale,
               "note": "one launch = one 32-frame batch of the 640x480 / 5 mm table sequence"},
              open(os.path.join(ROOT, "profiles", "integrate_kernel_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[1]
    for p in sys.argv[2:]:
        if p.endswith(".csv"):
            launches(p, tag)
        else:
            full(p, tag)
