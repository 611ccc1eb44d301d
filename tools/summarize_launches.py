"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v *= {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"<.*", "", name)[:80]
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print("total %.3f ms over %d launches" % (T / 1e6, sum(cnt.values())))
print("| share | ms | launches | kernel |\n|---|---|---|---|")
for n, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print("| %5.2f%% | %8.3f | %4d | %s |" % (100 * v / T, v / 1e6, cnt[n], n))
