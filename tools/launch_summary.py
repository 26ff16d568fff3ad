"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (time, launches, share)."""
import collections, csv, re, sys

def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        t *= {"us": 1e-3, "usecond": 1e-3, "ns": 1e-6, "nsecond": 1e-6}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        tot[name] += t; cnt[name] += 1
    s = sum(tot.values())
    print("total %.3f ms over %d launches" % (s, sum(cnt.values())))
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print("%-44s %4d  %9.3f ms  %5.1f%%" % (k[:44], cnt[k], v, 100 * v / s))

if __name__ == "__main__":
    main(sys.argv[1])
