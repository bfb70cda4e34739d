"""Record the DRAM traffic of one `ncu --set full` capture in profiles/traffic.json, keyed by the kernel source's hash
(bench.py reports it as roofline.traffic only while the source is unchanged).
usage: update_traffic.py <source.cu> <workload> <file.ncu-rep> <profiles/summary.txt>"""
import csv, hashlib, io, json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
source, workload, rep, capture = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
def metric(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
path = os.path.join(root, "profiles", "traffic.json")
rec = json.load(open(path))
sha = hashlib.sha256(open(os.path.join(root, "automative-rag_b200", "csrc", source), "rb").read()).hexdigest()[:16]
rec.setdefault(source, {})[workload] = {
    "dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "source_sha16": sha,
    "capture": capture, "how": "ncu --set full --clock-control none, one launch (scripts/profile_kernels.py)"}
json.dump(rec, open(path, "w"), indent=1)
print(source, workload, "read", rd, "write", wr, "sha", sha)
