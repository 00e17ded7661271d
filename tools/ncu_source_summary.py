"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump.

  python tools/ncu_source_summary.py dump.csv lines [N]   # hottest CUDA source lines
  python tools/ncu_source_summary.py dump.csv sass [pct]  # SASS in address order, rows >= pct %
"""
import csv
import sys
from collections import OrderedDict, defaultdict


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return 0.0


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, cur, line, text = None, None, "", ""
    sass = OrderedDict()
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        if r[0]:
            line, text = r[0], r[1].strip()
            continue
        if r[2].startswith("0x") and r[2] not in sass:
            d = dict(zip(hdr[4:], r[4:]))
            sass[r[2]] = dict(file=cur, line=line, text=text, sass=r[3].strip(),
                              inst=num(d["Instructions Executed"]), thr=num(d["Thread Instructions Executed"]),
                              samp=num(d["# Samples"]))
    return sass


def main():
    path, mode = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "lines"
    sass = load(path)
    tot = sum(v["inst"] for v in sass.values()) or 1.0
    thr = sum(v["thr"] for v in sass.values())
    samp = sum(v["samp"] for v in sass.values()) or 1.0
    print(f"warp-instructions {tot:.0f}  thread-instructions {thr:.0f}  mean active lanes {thr / tot:.1f}")
    if mode == "lines":
        top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
        agg = defaultdict(lambda: [0.0, 0.0, 0.0, ""])
        for v in sass.values():
            a = agg[(v["file"], v["line"])]
            a[0] += v["inst"]; a[1] += v["thr"]; a[2] += v["samp"]; a[3] = v["text"][:95]
        for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"{a[0] / tot * 100:5.1f}% inst {a[2] / samp * 100:5.1f}% smp  lanes {a[1] / max(a[0], 1):5.1f}  {f}:{l}  {a[3]}")
    else:
        pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.15
        for addr, v in sorted(sass.items(), key=lambda kv: int(kv[0], 16)):
            if v["inst"] / tot * 100 >= pct:
                print(f"{addr[-5:]} {v['inst'] / tot * 100:5.2f}% lanes {v['thr'] / max(v['inst'], 1):5.1f} smp {v['samp'] / samp * 100:5.2f}%  "
                      f"{v['file']}:{v['line']:>4}  {v['sass'][:70]}")


if __name__ == "__main__":
    main()
