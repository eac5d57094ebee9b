"""ncu per-launch CSV (gpu__time_duration.sum + DRAM bytes) -> the launch list of ONE step with each kernel's share.
usage: python scripts/launch_list.py launches.csv "header text" [first-kernel name fragment] > profiles/rNN_..._launch_list.txt
The step is delimited by the fused training kernel (tc_bwd_pair_kernel): the last complete step of the capture is used."""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
hdr = rows[0]
iid, ik, im, iu, iv = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
launches = {}
order = []
for r in rows[1:]:
    k = int(r[iid])
    if k not in launches:
        launches[k] = {"name": r[ik]}
        order.append(k)
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    if r[im] == "gpu__time_duration.sum":
        launches[k]["ms"] = v / 1e6 if u == "ns" else v / 1e3 if u in ("us", "usecond") else v
    else:
        scale = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1e-9)
        launches[k]["rd" if "read" in r[im] else "wr"] = v * scale
seq = [launches[k] for k in order]
fused = [i for i, l in enumerate(seq) if "tc_bwd_pair_kernel" in l["name"]]
# one step = from the first kernel after the previous step's last kernel to the last kernel before the next step's first
# prepack: cut at the pack_ctx launches that precede each fused kernel
first = sys.argv[3] if len(sys.argv) > 3 else "pack_ctx"          # name fragment of the first kernel of a step
packs = [i for i, l in enumerate(seq) if first in l["name"] and "unpack" not in l["name"]]
start = packs[-1]
end = len(seq)
prev_start = packs[-2] if len(packs) > 1 else 0
step = seq[prev_start:start] if len(packs) > 1 else seq[start:end]
tot = sum(l["ms"] for l in step)
print(sys.argv[2] if len(sys.argv) > 2 else "")
print(f"#   One step ({len(step)} launches, launch order): ms, DRAM read GB, write GB, share of the step's summed kernel time "
      f"({tot:.2f} ms under ncu: cold-cache, serialised)")
agg = {}
for l in step:
    name = re.sub(r"\(.*", "", l["name"]).replace("void ", "").replace("gloria::", "").replace("<unnamed>::", "")
    if name.startswith("at::"):
        name = "ATen " + re.sub(r"<.*", "", name.split("::")[-1])
    print(f"{l['ms']:11.4f} ms {l.get('rd', 0):9.3f} {l.get('wr', 0):9.3f} {100 * l['ms'] / tot:6.1f}%  {name[:100]}")
    a = agg.setdefault(name, [0.0, 0.0, 0.0, 0])
    a[0] += l["ms"]; a[1] += l.get("rd", 0); a[2] += l.get("wr", 0); a[3] += 1
print("# by kernel:")
for name, a in sorted(agg.items(), key=lambda t: -t[1][0]):
    print(f"#{a[0]:10.4f} ms {a[1]:9.3f} {a[2]:9.3f} {100 * a[0] / tot:6.1f}%  x{a[3]:<3d} {name[:100]}")
print(f"# DRAM total: read {sum(l.get('rd', 0) for l in step):.1f} GB, write {sum(l.get('wr', 0) for l in step):.1f} GB")
lib = [n for n in agg if n.startswith(("nvjet", "cutlass", "ATen", "cublas"))]
print("# library kernels in the step:", ", ".join(f"{n} ({agg[n][0]:.3f} ms)" for n in lib))
