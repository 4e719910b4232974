"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (shares of the step)."""
import collections
import csv
import re
import sys


def main(path, title=""):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")) / 1e3
        k = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    print("# %s\n" % title)
    print("Total %.1f us over %d launches (ncu serialises launches with cold caches: compare SHARES).\n" % (T, sum(cnt.values())))
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:36]:
        print("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k.replace("|", "\\|"), cnt[k], v, 100 * v / T, v / cnt[k]))


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
