"""Top stall sites from `ncu -i X.ncu-rep --page source --csv` output (SASS view): python tools/ncu_stalls.py file.csv [N]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    body = []
    for r in rows[hi + 1:]:
        if r and r[0] in ("Kernel Name", "Address"):   # the capture repeats per launch: keep the first
            break
        if len(r) >= len(hdr):
            body.append(r)
    total = sum(int(r[col["# Samples"]] or 0) for r in body)
    agg = {h: sum(int(r[col[h]] or 0) for r in body) for h in stall_cols}
    print(f"total samples {total}; instructions {len(body)}")
    print("by reason:", ", ".join(f"{k[6:]}={v} ({100 * v / max(total, 1):.1f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    body_idx = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:top]
    for i in sorted(body_idx):
        r = body[i]
        n = int(r[col["# Samples"]] or 0)
        reasons = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
        print(f"{i:5d} {n:6d} {100 * n / max(total, 1):5.1f}%  {r[col['Source']][:90]:90s} "
              + " ".join(f"{h}={v}" for v, h in reasons if v))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
