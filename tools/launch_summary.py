"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel launches / microseconds / share of the total."""
import csv
import re
import sys
from collections import defaultdict


def main(path, note=""):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k_i, v_i, u_i = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = re.sub(r"^(void )?(hkcsa::)?", "", r[k_i])
        name = re.sub(r"\(.*", "", name)
        ns = float(r[v_i].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[u_i], 1.0)
        tot[name][0] += 1
        tot[name][1] += ns / 1e3
    total = sum(v[1] for v in tot.values())
    if note:
        print(note)
    print(f"{len(rows) - 1} launches captured, gpu__time_duration summed per kernel, serialised and cold-cache: compare SHARES.")
    print(f"{'kernel':70s} {'launches':>8s} {'us':>12s} {'share':>7s}")
    for name, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:70]:70s} {c:8d} {us:12.1f} {us / total:7.3f}")


if __name__ == "__main__":
    main(*sys.argv[1:3])
