"""Per-source-line stall samples of one kernel of an .ncu-rep (needs -lineinfo and --import-source on):
    python profiles/scripts/ncu_source_lines.py rep.ncu-rep <launch index> [top N]"""
import csv
import io
import subprocess
import sys

rep, skip = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
name = [r[1] for r in rows if r and r[0] == "Function Name"]
hdr = next(r for r in rows if r and r[0] == "Line No")
i_s, i_i = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
cur_file, lines = "", []
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    if len(r) > i_i and r[0].isdigit() and r[2] == "-":  # a source line (its SASS rows follow, with an address)
        lines.append((int(r[i_s] or 0), int(r[i_i] or 0), cur_file, int(r[0]), r[1].strip()))
tot = sum(l[0] for l in lines)
print(name[0] if name else "?", "- total samples", tot)
for s, i, f, ln, src in sorted(lines, reverse=True)[:top]:
    print("%6d %5.1f%%  inst %9d  %s:%d: %s" % (s, 100.0 * s / max(tot, 1), i, f, ln, src[:100]))
