"""Print per-kernel headline metrics and the top stalled SASS lines of an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else None; ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_read.sum", "sm__cycles_elapsed.max"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
for r in rows[2:]:
    if pat and pat not in r[hdr.index("Kernel Name")]:
        continue
    print(" | ".join(f"{w.split('.')[0][-26:]}={r[i][:70]}" for w, i in idx))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, cur, h = [], None, None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; kern.append(cur); h = None; continue
    if r and r[0] == "Address":
        h = r; cur["hdr"] = r; continue
    if cur is not None and h and r:
        cur["rows"].append(r)
seen = set()
for k in kern:
    if pat and pat not in k["name"]:
        continue
    if k["name"] in seen:
        continue
    seen.add(k["name"])
    h = k["hdr"]; iS = h.index("# Samples"); isrc = h.index("Source")
    sc = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[iS] or 0) for r in k["rows"])
    agg = sorted(((h[i], sum(int(r[i] or 0) for r in k["rows"])) for i in sc), key=lambda x: -x[1])[:6]
    print("\n==", k["name"][:120], "samples", tot, agg)
    for r in sorted(k["rows"], key=lambda r: -int(r[iS] or 0))[:ntop]:
        st = sorted(((h[i], int(r[i] or 0)) for i in sc if int(r[i] or 0) > 0), key=lambda x: -x[1])[:2]
        print(r[iS].rjust(6), r[isrc][:84].ljust(84), st)
