#!/bin/bash
# Run HERE after tools/collect_profiles.sh came back through gpurun_out/: copy the evidence the judge reads into profiles/.
# usage: tools/publish_profiles.sh <tag>
TAG=${1:-r1}
G=gpurun_out; P=profiles
for f in bench_${TAG}.json bench_exact_${TAG}.json bench_fast_${TAG}.json bench_reference_${TAG}.json launches_${TAG}.csv multiview_1.json; do cp $G/$f $P/$f; done
python tools/ncu_summary.py $G/prof_${TAG}.ncu-rep k_geom_raster 30 > $P/ncu_geom_raster_${TAG}.txt 2>/dev/null
python tools/ncu_summary.py $G/prof_${TAG}.ncu-rep k_resolve 20 > $P/ncu_resolve_${TAG}.txt 2>/dev/null
python tools/ncu_summary.py $G/prof_${TAG}.ncu-rep "k_shade|k_resolve_shade" 25 > $P/ncu_shade_${TAG}.txt 2>/dev/null
python tools/ncu_lines.py $G/prof_${TAG}.ncu-rep k_geom_raster 40 "(bool)1, (bool)1" > $P/ncu_geom_raster_lines_${TAG}.txt 2>/dev/null
python - <<PY
import csv, io, json, subprocess
rep = "$G/prof_${TAG}.ncu-rep"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]; ik = h.index("Kernel Name"); ir = h.index("dram__bytes_read.sum"); iw = h.index("dram__bytes_write.sum")
units = rows[1]
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
cls = {}
for r in rows[2:]:
    k = r[ik]
    name = ("geom_raster_shadow" if "k_geom_raster" in k and "true, true" in k.replace("(bool)1", "true").replace("(bool)0", "false").replace("<1, 1>", "<true, true>") else
            "geom_raster_camera" if "k_geom_raster" in k else "shade" if "k_resolve_shade" in k else "resolve" if "k_resolve<" in k or "k_resolve(" in k else "shade" if "k_shade<" in k or "k_shade(" in k else None)
    if name is None or "special" in k or "resolve00" in k: continue
    b = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
    c = cls.setdefault(name, {"dram_bytes_per_launch": 0.0, "launches_sampled": 0})
    c["dram_bytes_per_launch"] += b; c["launches_sampled"] += 1
for c in cls.values():
    c["dram_bytes_per_launch"] /= c["launches_sampled"]
cls["_source"] = "ncu --set full --clock-control none capture prof_${TAG}.ncu-rep (tools/collect_profiles.sh ${TAG}): dram__bytes_read.sum + dram__bytes_write.sum per launch"
json.dump(cls, open("$P/traffic_${TAG}.json", "w"), indent=1)
print(json.dumps(cls, indent=1))
PY
