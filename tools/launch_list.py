"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel table kept under profiles/.
Usage: python tools/launch_list.py launches.csv "header comment" [more header lines...] > profiles/rN_launches_X.txt"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in rows:
        name = re.sub(r"\(.*", "", r[4].replace("void ", ""))
        name = re.sub(r"<\(bool\)(\d)>", r"<\1>", name)
        ns = float(r[14])
        if r[13] == "us":
            ns *= 1e3
        elif r[13] == "ms":
            ns *= 1e6
        tot[name] += ns
        cnt[name] += 1
    for h in sys.argv[2:]:
        print("# " + h)
    print("# cold-cache, serialised: compare SHARES.")
    print(f"# {'kernel':42} {'launches':>9} {'total us':>12} {'share':>7} {'us/launch':>11}")
    total = sum(tot.values())
    for name, ns in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{name:44} {cnt[name]:>9} {ns / 1e3:>12.1f} {100 * ns / total:>6.1f}% {ns / 1e3 / cnt[name]:>11.1f}")
    if "k_verify_half_main<0>" in tot:
        a, b = tot["k_verify_half_prep<0>"], tot["k_verify_half_main<0>"]
        print(f"# step = k_verify_half_prep + k_verify_half_main: main share of the step = {100 * b / (a + b):.1f}%, prep {100 * a / (a + b):.1f}%")


if __name__ == "__main__":
    main()
