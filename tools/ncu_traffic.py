"""DRAM traffic of one verify step from an `ncu --page raw --csv` export: dram__bytes_read.sum + dram__bytes_write.sum of the
first launch of every kernel of the step (k_verify_half_prep, k_half_sort_*, k_verify_half_main).  bench.py reports the sum as
roofline.traffic.  Usage: python tools/ncu_traffic.py raw.csv [log2n] > profiles/r2_verify_traffic.json"""
import csv
import json
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].strip()
        if name in per:
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[ix[k]].replace(",", "")) * UNIT[units[ix[k]]]
        per[name] = int(tot)
    print(json.dumps({"log2n": log2n, "dram_bytes_per_step": sum(per.values()), "per_kernel": per,
                      "source": "profiles/r2_ncu_verify.txt: dram__bytes_read.sum + dram__bytes_write.sum of the launches of one 2^%d-signature step, "
                                "ncu --set full --clock-control none on the final round-2 build (tools/gpu_final3_n1.sh, tools/ncu_traffic.py)" % log2n}, indent=1))


if __name__ == "__main__":
    main()
