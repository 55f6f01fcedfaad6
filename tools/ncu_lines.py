"""Per-source-line summary of an ncu --set full --import-source on capture (source page, cuda+sass view):
warp-stall samples, executed warp instructions and the dominant stall reasons of every hot line.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n] > profiles/rNN_x_lines.txt
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    lines = {}
    cur_file, hdr, func = None, None, None
    total_samples = total_inst = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_samples, i_inst = hdr.index("# Samples"), hdr.index("Instructions Executed")
            stall_cols = [(i, h[len("stall_"):]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or not r[0].strip().isdigit():
            continue                                  # SASS rows carry no line number
        if len(r) > len(hdr):
            # a source line with unescaped quotes (inline asm) splits into extra cells: the counters are the LAST cells
            extra = len(r) - len(hdr)
            r = [r[0], ",".join(r[1:2 + extra])] + r[2 + extra:]
        try:
            samples, inst = int(float(r[i_samples] or 0)), int(float(r[i_inst] or 0))
        except (ValueError, IndexError):
            continue
        if samples == 0 and inst == 0:
            continue
        key = (cur_file, int(r[0]))
        e = lines.setdefault(key, {"samples": 0, "inst": 0, "src": r[1].strip(), "stalls": {}})
        e["samples"] += samples
        e["inst"] += inst
        for i, name in stall_cols:
            try:
                v = int(float(r[i] or 0))
            except ValueError:
                v = 0
            if v:
                e["stalls"][name] = e["stalls"].get(name, 0) + v
        total_samples += samples
        total_inst += inst
    print("# %s: %d warp-stall samples, %d warp instructions executed" % (func, total_samples, total_inst))
    print("# samples   share   instructions  file:line  top stall reasons | source")
    for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = ",".join("%s:%d" % kv for kv in sorted(e["stalls"].items(), key=lambda kv: -kv[1])[:3])
        print("%8d  %5.1f%%  %12d  %s:%d  %s | %s" % (e["samples"], 100.0 * e["samples"] / max(total_samples, 1), e["inst"], f, ln, st, e["src"][:110]))
    # by function-sized regions: cumulative share per file
    per_file = {}
    for (f, ln), e in lines.items():
        a = per_file.setdefault(f, [0, 0])
        a[0] += e["samples"]
        a[1] += e["inst"]
    for f, (s_, i_) in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
        print("# file %-28s samples %6d (%4.1f%%)  instructions %10d (%4.1f%%)" % (f, s_, 100.0 * s_ / max(total_samples, 1), i_, 100.0 * i_ / max(total_inst, 1)))


if __name__ == "__main__":
    main()
