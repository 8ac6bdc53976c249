"""Condenses the ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1a.csv profiles/r1a_step_launches.csv
    python profiles/summarize.py kernel   gpurun_out/prof_aggfwd_r1d.ncu-rep profiles/r1d_agg_fwd.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def launches(src, dst):
    rows = []
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    for row in csv.DictReader(lines):
        rows.append((row["Kernel Name"], float(row["Metric Value"].replace(",", ""))))
    idx = [i for i, r in enumerate(rows) if "plan_count_kernel" in r[0]]
    pairs = list(zip(idx[:-1], idx[1:]))
    # one full steady-state step (plan rebuild + forward/backward/Adam): the most common segment length, which skips
    # the first step (lazy optimizer-state initialisation) and the short plan-only segments of the e2e phase
    lens = collections.Counter(p[1] - p[0] for p in pairs if p[1] - p[0] > 50)
    common = lens.most_common(1)[0][0]
    a, b = [p for p in pairs if p[1] - p[0] == common][0]
    seg = rows[a:b]
    tot = sum(v for _, v in seg)
    agg = collections.OrderedDict()
    for n, v in seg:
        n = re.sub(r"\(.*", "", n)[:100]
        agg.setdefault(n, [0.0, 0])
        agg[n][0] += v
        agg[n][1] += 1
    with open(dst, "w") as f:
        f.write("# one training step under ncu --metrics gpu__time_duration.sum (serialised, cold cache): %d kernels, "
                "%.1f us total\n" % (len(seg), tot / 1000))
        f.write("kernel,launches,total_us,share_pct\n")
        for n, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
            f.write('"%s",%d,%.1f,%.2f\n' % (n, c, v / 1000, 100 * v / tot))


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    out = []
    for row in r[2:]:
        d = dict(zip(hdr, row))
        out.append("kernel: %s" % d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                out.append("  %-85s %s %s" % (k, d[k], units[hdr.index(k)]))
    srcp = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(srcp.splitlines()))
    h, data = None, []
    for rr in rows:
        if rr and rr[0] == "Address":
            if h is not None:
                break
            h = rr
            continue
        if h is not None and len(rr) == len(h):
            data.append(rr)
    if h:
        ix, so = h.index("Instructions Executed"), h.index("Source")
        tot = sum(int(x[ix]) for x in data)
        c = collections.Counter()
        for x in data:
            t = x[so].split()
            op = t[1] if t[0].startswith("@") else t[0]
            c[op.split(".")[0]] += int(x[ix])
        out.append("  SASS instructions in kernel: %d; executed warp instructions: %d" % (len(data), tot))
        out.append("  opcode mix: " + ", ".join("%s %.1f%%" % (o, 100 * v / tot) for o, v in c.most_common(14)))
    open(dst, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
