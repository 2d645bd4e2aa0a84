"""Box-side: turns an .ncu-rep into two small CSV files (the report itself can exceed what gpurun brings back):
<out>_raw.csv (every metric of every captured launch) and <out>_source.csv (the `top` hottest SASS lines by samples;
top = 0: every line in address order, gzip-compressed, <out>_source_all.csv.gz)."""
import csv
import gzip
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 400
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(out + "_raw.csv", "w").write(raw)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr_i = next((i for i, r in enumerate(rows) if r and r[0] == "Address"), None)
if hdr_i is not None:
    hdr = rows[hdr_i]
    data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    col = hdr.index("# Samples")
    if top == 0:
        keep = [i for i, h in enumerate(hdr) if h in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed",
                                                        "Warp Stall Sampling (All Samples)", "Warp Stall Sampling (Not-issued Samples)")
                or (h.startswith("stall_") and "Not Issued" not in h)]
        with gzip.open(out + "_source_all.csv.gz", "wt", newline="") as f:
            w = csv.writer(f)
            w.writerow([hdr[i] for i in keep])
            for r in data:
                w.writerow([r[i] for i in keep])
        top = 400
    data.sort(key=lambda r: -int(r[col] or 0))
    with open(out + "_source.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(hdr + ["total_samples", "total_lines"])
        tot = sum(int(r[col] or 0) for r in data)
        for r in data[:top]:
            w.writerow(r + [tot, len(data)])
print("exported", rep)
