"""Top CUDA source lines by warp-stall samples in an .ncu-rep (needs -lineinfo + --import-source on).
Usage: python tools/ncu_src.py rep.ncu-rep [kernel-substring] [ntop]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else None; ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fn, h, rows, done = None, None, [], set()
def flush():
    if fn and rows and (not pat or pat in fn) and fn not in done:
        done.add(fn)
        iS = h.index("# Samples"); isrc = 1; il = h.index("Line No"); ie = h.index("Instructions Executed")
        tot = sum(int(r[iS] or 0) for r in rows)
        print("==", fn[:100], "samples", tot)
        for r in sorted(rows, key=lambda r: -int(r[iS] or 0))[:ntop]:
            print(r[iS].rjust(6), ("%5.1f%%" % (100.0 * int(r[iS] or 0) / max(tot, 1))), ("L" + r[il]).ljust(6), r[ie].rjust(9), r[isrc].strip()[:130])
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "Function Name":
        flush(); fn, h, rows = r[1], None, []
    elif r[0] == "Line No": h = r
    elif h and r[0].isdigit() and len(r) > 6 and r[2] == "-": rows.append(r)
flush()
