"""Summarise the warp-stall samples of every kernel in an .ncu-rep (captured with --set full --import-source on):
per kernel the stall-reason mix and the SASS lines that collect the most samples.

    python tools/ncu_stalls.py report.ncu-rep out.txt [top]
"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, which, extra=()):
    r = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv", *extra], capture_output=True, text=True)
    return list(csv.reader(io.StringIO(r.stdout)))


def main():
    rep, out, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 14
    raw = page(rep, "raw")
    hd = raw[0]
    dur = [r[hd.index("gpu__time_duration.sum")] for r in raw[2:]]
    stall_cols = [(i, h) for i, h in enumerate(hd) if "warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    src = page(rep, "source", ("--print-source", "sass"))
    starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    sections = []
    for k, s0 in enumerate(starts):
        h = src[s0 + 1]
        si, ci = h.index("Source"), h.index("# Samples")
        end = starts[k + 1] if k + 1 < len(starts) else len(src)
        rows = []
        for r in src[s0 + 2:end]:
            try:
                rows.append((int(r[ci]), r[si]))
            except (ValueError, IndexError):
                pass
        sections.append((src[s0][1], rows))
    if len(sections) == 2 * len(dur):          # every launch is listed twice (two source views): keep one
        sections = sections[::2]
    with open(out, "w") as f:
        for k, (nm, rows) in enumerate(sections):
            tot = sum(v for v, _ in rows) or 1
            f.write(f"kernel {nm[:90]}  {dur[k] if k < len(dur) else '?'} us  samples {tot}\n")
            if k < len(raw) - 2:
                r = raw[2 + k]
                st = sorted(((float(r[i].replace(',', '')), h2.split("issue_stalled_")[1].split("_per_issue")[0]) for i, h2 in stall_cols
                             if r[i] not in ("", "n/a")), reverse=True)[:8]
                ssum = sum(v for v, _ in st) or 1
                f.write("  stall mix: " + "  ".join(f"{n} {100 * v / ssum:.0f}%" for v, n in st) + "\n")
            for v, s in sorted(rows, reverse=True)[:top]:
                f.write(f"  {100 * v / tot:6.2f}%  {s.strip()[:120]}\n")
            f.write("\n")


if __name__ == "__main__":
    main()
