"""Condense ncu output into the small text summaries that are committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv            # per-kernel launch count / time / share
    python tools/ncu_summary.py full     gpurun_out/pool_full.ncu-rep       # per-launch key counters of a --set full capture
    python tools/ncu_summary.py stalls   gpurun_out/pool_full.ncu-rep K     # top stall sites (source page) of kernel K
"""
import collections
import csv
import io
import subprocess
import sys

EXACT = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]
SUBSTR = ["pipe_tensor", "pipe_tc", "pipe_tmem", "inst_executed_pipe_uniform"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def full(rep):
    hdr, units, rows = raw_rows(rep)
    name_i = hdr.index("Kernel Name")
    for r in rows:
        print("== %s" % r[name_i].split("(")[0])
        for i, h in enumerate(hdr):
            extra = (any(s in h for s in SUBSTR) and ".avg." in h and h.endswith("pct_of_peak_sustained_elapsed")
                     and "Triage" not in h and r[i] not in ("", "0", "0.000000"))
            if h in EXACT or extra:
                print("   %-86s %-8s %s" % (h, units[i], r[i]))
        # stall breakdown: warp-cycles per issued instruction by reason
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("   stall reasons (warps stalled per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:7]))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}[d["Metric Unit"]]
        k = d["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(k, [0, 0.0, d["Grid Size"], d["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("%-26s %7s %12s %11s %7s  %s" % ("kernel", "launches", "total ms", "avg us", "share", "grid x block (first launch)"))
    for k, (n, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-26s %7d %12.3f %11.2f %6.1f%%  %s x %s" % (k[:26], n, t / 1e6, t / n / 1e3, 100 * t / tot, g, b))
    print("%-26s %7d %12.3f" % ("total", sum(v[0] for v in agg.values()), tot / 1e6))


def stalls(rep, kernel, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kernel, "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    items = []
    for r in rows:
        if "Source" in r and any("Sampling" in c for c in r):
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            key = [c for c in hdr if c.startswith("Warp Stall Sampling (All")]
            try:
                items.append((int(d[key[0]]), d.get("Address", ""), d["Source"]))
            except (ValueError, IndexError, KeyError):
                pass
    tot = sum(i[0] for i in items) or 1
    for n, addr, src in sorted(items, reverse=True)[:top]:
        print("%6d %5.1f%%  %s  %s" % (n, 100.0 * n / tot, addr, src[:120]))




def lines(rep, kernel, cubin, top=40):
    """Attribute warp-stall samples of `kernel` to source lines: ncu SASS page joined with `nvdisasm -g` of `cubin`."""
    import re
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
    off2line, cur, in_fn = {}, None, False
    for ln in dis:
        if ln.startswith(".text."):
            in_fn = kernel in ln
            continue
        if not in_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m and cur:
            off2line[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kernel, "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    agg = collections.Counter()
    reasons = collections.defaultdict(collections.Counter)
    hdr, base, launches_seen = None, None, 0
    for r in rows:
        if r and r[0] == "Address":
            hdr, base = r, None
            launches_seen += 1
            continue
        if hdr is None or len(r) != len(hdr) or launches_seen > 1:
            continue
        d = dict(zip(hdr, r))
        addr = int(d["Address"], 16)
        base = addr if base is None else base
        n = int(d["Warp Stall Sampling (All Samples)"] or 0)
        key = off2line.get(addr - base, ("?", 0))
        agg[key] += n
        for c in hdr:
            if c.startswith("stall_") and "Not Issued" not in c and d[c] not in ("", "0"):
                reasons[key][c[6:]] += int(d[c])
    tot = sum(agg.values()) or 1
    src_cache = {}
    for (f, l), n in agg.most_common(top):
        text = ""
        for cand in ("ataxxzero_b200/csrc/" + f,):
            try:
                src_cache.setdefault(cand, open(cand).read().splitlines())
                text = src_cache[cand][l - 1].strip()
            except (OSError, IndexError):
                pass
        why = ",".join("%s:%d" % kv for kv in reasons[(f, l)].most_common(2))
        print("%6d %5.1f%%  %-14s:%-4d %-28s %s" % (n, 100.0 * n / tot, f, l, why, text[:90]))


if __name__ == "__main__":
    {"full": full, "launches": launches, "stalls": stalls, "lines": lines}[sys.argv[1]](*sys.argv[2:])
