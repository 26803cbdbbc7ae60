"""Summarise an `ncu --set full` capture (one launch per kernel, tools/kbench.py under KBENCH_PROFILE=1) into
  profiles/<name>.txt        one line per launch: duration, DRAM bytes read / written, tensor-pipe %, warps active %, registers, grid
  profiles/ncu_traffic.json  {key: {kernel, dram_bytes_read, dram_bytes_write, source}} for the kernels bench.py reports a roofline for
Usage: python tools/ncu_traffic.py gpurun_out/r02_kernels.ncu-rep profiles/r02_kernels_full.txt"""
import csv, io, json, os, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def val(r, name, scale_bytes=False):
    if name not in col: return None
    v = r[col[name]].replace(",", "")
    try: x = float(v)
    except ValueError: return None
    if scale_bytes:
        u = units[col[name]].lower()
        x *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return x
def dur_us(r):
    x = val(r, "gpu__time_duration.sum"); u = units[col["gpu__time_duration.sum"]].lower()
    return x * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1)
lines, traffic = [], {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    rd, wr = val(r, "dram__bytes_read.sum", True), val(r, "dram__bytes_write.sum", True)
    t = dur_us(r)
    lines.append(f"{t:9.1f} us  dram rd {rd/1e6:8.2f} MB wr {wr/1e6:8.2f} MB ({(rd+wr)/t/1e3:7.1f} GB/s)  tensor {val(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') or 0:5.1f}%  "
                 f"warps {val(r,'sm__warps_active.avg.pct_of_peak_sustained_active') or 0:5.1f}%  regs {int(val(r,'launch__registers_per_thread') or 0):3d}  grid {int(val(r,'launch__grid_size') or 0):6d}  {name[:150]}")
    key = None
    if "gemm_ws2_kernel<1, 1>" in name.replace("(bool)", "").replace("(int)", "") and "ffn_up_silu" not in traffic: key = "ffn_up_silu"
    if "layernorm_bwd_kernel" in name and "layernorm_bwd" not in traffic: key = "layernorm_bwd"
    if key: traffic[key] = {"kernel": name[:120], "dram_bytes_read": rd, "dram_bytes_write": wr, "us_under_ncu": round(t, 2), "source": os.path.relpath(out)}
open(out, "w").write("# ncu --set full --clock-control none, one launch per kernel (tools/kbench.py, KBENCH_PROFILE=1, L2 flushed before the launch); cold-cache, serialised times\n" + "\n".join(lines) + "\n")
tj = os.path.join(os.path.dirname(out), "ncu_traffic.json")
old = json.load(open(tj)) if os.path.exists(tj) else {}
old.update(traffic)
json.dump(old, open(tj, "w"), indent=1)
print("\n".join(lines)); print(traffic)
