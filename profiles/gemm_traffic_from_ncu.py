"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch, GEMM kernels of one training
step) -> profiles/r01_gemm_traffic.json, which bench.py reads to fill roofline.traffic.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:gemm_tcgen05 --csv --log-file gpurun_out/gemm_dram.csv python bench.py --steps 1 --warmup 1 --eager ...
    python profiles/gemm_traffic_from_ncu.py gpurun_out/gemm_dram.csv 601 [profiles/r02_gemm_traffic.json]
(round 2: the capture comes from profiles/one_step.py under `ncu --profile-from-start off`, i.e. exactly one step)
"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 601
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
iid, name, unit, val, met = (hdr.index(k) for k in ("ID", "Metric Name", "Metric Unit", "Metric Value", "Metric Name"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
launch = {}
for r in rows[h + 1:]:
    if len(r) <= val:
        continue
    try:
        v = float(r[val].replace(",", "")) * scale.get(r[unit], 1.0)
    except ValueError:
        continue
    launch.setdefault(r[iid], {})[r[name]] = v
ids = sorted(launch, key=lambda k: int(k))
ids = ids[-per_step:]  # the last step captured
rd = sum(launch[i].get("dram__bytes_read.sum", 0.0) for i in ids)
wr = sum(launch[i].get("dram__bytes_write.sum", 0.0) for i in ids)
us = sum(launch[i].get("gpu__time_duration.sum", 0.0) for i in ids)
out = {"launches": len(ids), "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_launch": (rd + wr) / max(len(ids), 1),
       "kernel_us_under_ncu": us,
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (cache control: flush between launches)"}
json.dump(out, open(sys.argv[3] if len(sys.argv) > 3 else "profiles/r01_gemm_traffic.json", "w"), indent=1)
print(json.dumps(out))
