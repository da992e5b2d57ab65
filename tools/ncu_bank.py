"""Development: SASS lines of an .ncu-rep with the most excessive shared-memory wavefronts (bank conflicts), with source file:line."""
import csv, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"] if len(sys.argv) > 3 else
                     ["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ce, cw, ci, cs, cx = (hdr.index(n) for n in ("L1 Wavefronts Shared Excessive", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "Source", "Instructions Executed"))
data = []
for r in rows[hi + 1:]:
    try:
        data.append((float(r[ce] or 0), float(r[cw] or 0), float(r[ci] or 0), r[cs].strip(), r[cx]))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print(f"excessive shared wavefronts total {tot:.0f}; all shared wavefronts {sum(d[1] for d in data):.0f}")
for e, w, i, s_, x in sorted(data, key=lambda t: -t[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"  {100*e/max(tot,1):5.1f}%  excess={e:12.0f} total={w:12.0f} ideal={i:12.0f} exec={x:>10s}  {s_[:80]}")
