"""Instruction mix / stall samples per SASS opcode from `ncu --page source --csv` output."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
idx = {k: i for i, k in enumerate(rows[h])}
ops, samples, tot = collections.Counter(), collections.Counter(), 0
for r in rows[h + 1:]:
    if len(r) < len(idx) or not r[idx["Instructions Executed"]].isdigit():
        continue
    sass = r[idx["Source"]].strip()
    n = int(r[idx["Instructions Executed"]])
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    op = m.group(2).split(".")[0] if m else sass[:10]
    ops[op] += n; tot += n
    samples[op] += int(r[idx["# Samples"]] or 0)
S = sum(samples.values())
print("total warp instructions", tot, "samples", S)
for o, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print("%-12s %12d %5.1f%%   stall samples %5.1f%%" % (o, n, 100 * n / tot, 100 * samples[o] / max(S, 1)))
