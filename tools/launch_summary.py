"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("eec::<unnamed>::", "eec::")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("nsecond", "ns") else v * 1e3 if u in ("msecond", "ms") else v
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print(f"total {T/1e3:.3f} ms over {sum(cnt.values())} launches ({path})")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v:10.1f} us {100*v/T:5.1f}%  n={cnt[k]:4d}  avg={v/cnt[k]:8.1f} us  {k[:100]}")
