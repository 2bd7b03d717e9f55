"""Top stall locations of one kernel from an `ncu --set full --import-source on` report (run where the report lives; the raw report is too big
to ship).   python tools/ncu_source_top.py report.ncu-rep [N=45]   -> SASS lines sorted by warp-stall samples, with their main stall reasons"""
import csv
import io
import subprocess
import sys

rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1] if rows and len(rows[0]) > 1 else "?")
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("instructions %d, samples %d; by reason: %s" % (len(data), tot, ", ".join("%s %d" % kv for kv in sorted(agg.items(), key=lambda kv: -kv[1])[:8])))
for i, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][ix["# Samples"]] or 0))[:n]:
    s = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print("%6d %5.1f%%  #%-5d %-64s %s" % (s, 100.0 * s / max(tot, 1), i, r[ix["Source"]][:64], ", ".join("%s %d" % (k, v) for v, k in st if v)))
print("---- every synchronisation / tensor / TMA instruction with samples (address order)")
for i, r in enumerate(data):
    src = r[ix["Source"]]
    s = int(r[ix["# Samples"]] or 0)
    if s and any(k in src for k in ("SYNCS", "UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "BAR.", "LDTM", "STTM", "UBLKCP", "DEPBAR", "BRA")):
        st = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
        print("%6d %5.1f%%  #%-5d %-64s %s" % (s, 100.0 * s / max(tot, 1), i, src[:64], ", ".join("%s %d" % (k, v) for v, k in st if v)))
