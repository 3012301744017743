"""Per-kernel share of the solver launches in an ncu launch list (csv of --metrics gpu__time_duration.sum).

usage: python profiles/launch_shares.py profiles/r01_launches.csv [first_kernel_substring]
Only the launches from the first solver kernel (default: the first t_init_be / k_init) onwards are counted,
so the one-off mesh/assembly kernels of the set-up do not dilute the shares of the time loop.
"""
import csv
import re
import sys
from collections import defaultdict


def main(path, start="init"):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for rec in csv.DictReader(lines):
        name = re.sub(r"^void ", "", rec["Kernel Name"]).split("(")[0].split("<")[0]
        rows.append((name, float(rec["Metric Value"]) / 1e3))
    first = next((i for i, (n, _) in enumerate(rows) if start in n), 0)
    tot, cnt = defaultdict(float), defaultdict(int)
    for n, us in rows[first:]:
        tot[n] += us
        cnt[n] += 1
    total = sum(tot.values())
    print(f"# {path}: {len(rows) - first} launches from the first '{start}' kernel on, {total:.1f} us")
    for n in sorted(tot, key=tot.get, reverse=True):
        print(f"{n:30s} n={cnt[n]:4d} total={tot[n]:9.1f} us share={100 * tot[n] / total:5.1f}% avg={tot[n] / cnt[n]:7.1f} us")


if __name__ == "__main__":
    main(*sys.argv[1:])
