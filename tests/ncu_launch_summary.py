"""Summarises an ncu launch list of one steady-state step into the two files bench.py / profiles/README.md refer to.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/launches.csv python tests/profile_step.py 32 [enc] [fa]
    python tests/ncu_launch_summary.py gpurun_out/launches.csv profiles/r2_ 32

writes <prefix>launch_shares.txt (per-kernel cold-cache time share and DRAM bytes) and <prefix>conv_dram_bytes.json
(dram__bytes_read + write of all convolution launches of the step = bench.py's roofline.traffic)."""
import collections
import csv
import json
import re
import sys

src, prefix, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
launches = collections.OrderedDict()   # ncu launch id -> {"kernel", metric: value}
for row in csv.reader(open(src)):
    if len(row) < 15 or not row[0].isdigit():
        continue
    d = launches.setdefault(row[0], {"kernel": row[4]})
    d[row[12]] = float(row[14].replace(",", ""))


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    return name.replace("tsp::", "")


agg = collections.OrderedDict()
for d in launches.values():
    a = agg.setdefault(short(d["kernel"]), [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0) / 1e6          # ns -> ms
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
total = sum(a[1] for a in agg.values())
lines = [f"all kernels {total:.3f} ms cold-cache serialised, {len(launches)} launches (one steady-state step, {batch} clips)"]
for k, (n, ms, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{ms:8.3f} ms {100 * ms / total:5.1f}% n={n:3d} read {rd / 1e9:7.2f} GB write {wr / 1e9:7.2f} GB  {k}")
open(prefix + "launch_shares.txt", "w").write("\n".join(lines) + "\n")
conv = [a for k, a in agg.items() if k.startswith("conv_")]
out = {"batch_clips": batch, "dram_bytes_per_step": int(sum(a[2] + a[3] for a in conv)),
       "dram_read_bytes": int(sum(a[2] for a in conv)), "dram_write_bytes": int(sum(a[3] for a in conv)),
       "conv_launches": int(sum(a[0] for a in conv)), "conv_time_ms_cold": round(sum(a[1] for a in conv), 3),
       "all_kernels_time_ms_cold": round(total, 3),
       "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                 "--profile-from-start off python tests/profile_step.py %d (B200, one steady-state step; %s)" % (batch, src)}
json.dump(out, open(prefix + "conv_dram_bytes.json", "w"), indent=1)
print("\n".join(lines))
print(json.dumps(out))
