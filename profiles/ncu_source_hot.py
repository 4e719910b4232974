"""Hot source lines of one kernel from an ncu report: `python profiles/ncu_source_hot.py report.ncu-rep kernel_regex [n]`.
Runs `ncu -i ... --page source --csv --print-source cuda,sass` and sums the warp-stall samples per (file, line)."""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + rx, "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fpath, hdr, acc = None, None, {}
stall_cols = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or r[0] in ("Function Name",) or not r[0].isdigit():
        continue
    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    n = num(r[i_s])
    if n == 0:
        continue
    st = {h: num(r[i]) for i, h in stall_cols if num(r[i])}
    k = (fpath, int(r[0]), r[1].strip()[:110])
    a = acc.setdefault(k, [0, {}])
    a[0] += n
    for h, v in st.items():
        a[1][h] = a[1].get(h, 0) + v
tot = sum(v[0] for v in acc.values())
print("total samples", tot)
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    main = ", ".join("%s %d" % (h[6:], c) for h, c in sorted(v[1].items(), key=lambda x: -x[1])[:3])
    print("%5.1f%%  %s:%d  %s   [%s]" % (100.0 * v[0] / tot, k[0], k[1], k[2], main))
