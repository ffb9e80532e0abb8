"""Per-kernel SASS census from an `ncu --page source --csv` dump: executed warp instructions by opcode class, normalised
per butterfly when the butterfly count of the launch is given.
usage: python tools/sass_census.py <src.csv> [--kernel SUBSTR] [--butterflies N] [--top 25] [--stalls]"""
import argparse, collections, csv, re, sys

csv.field_size_limit(10 ** 9)
ap = argparse.ArgumentParser()
ap.add_argument("src")
ap.add_argument("--kernel", default=None)
ap.add_argument("--butterflies", type=float, default=0, help="butterflies per launch (rows * N/2 * logN): adds a per-butterfly column")
ap.add_argument("--top", type=int, default=30)
ap.add_argument("--stalls", action="store_true", help="also list the instructions with the most stall samples")
a = ap.parse_args()

CLASSES = [
    ("IMAD.WIDE (64-bit product)", r"^IMAD\.WIDE"), ("IMAD.HI", r"^IMAD\.HI"), ("IMAD (32-bit mul-add)", r"^IMAD(?!\.MOV|\.WIDE|\.HI|\.SHL|\.IADD)"),
    ("IMAD.MOV/SHL/IADD (moves on the FMA pipe)", r"^IMAD\.(MOV|SHL|IADD)"), ("IADD3 / IADD.64 / LEA", r"^(IADD|LEA|UIADD|ULEA)"), ("MOV / PRMT / SEL", r"^(MOV|PRMT|SEL|UMOV|USEL)"),
    ("LOP3 / SHF", r"^(LOP3|SHF|ULOP|USHF)"), ("ISETP / PLOP", r"^(ISETP|PLOP|UISETP|UPLOP)"),
    ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LD/ST shared::cluster", r"^(LD|ST)\.(E\.)?.*|^(LDSM|MAPA|UCGABAR|CGABAR)"),
    ("BAR / WARPSYNC / barrier.cluster", r"^(BAR|WARPSYNC|BSYNC|BSSY|UCGABAR|CGAERRBAR|ERRBAR|MEMBAR|FENCE|ACQBULK)"), ("BRA / EXIT / control", r"^(BRA|EXIT|CALL|RET|NOP|BREAK|YIELD)"),
    ("LDC / S2R / S2UR / uniform", r"^(LDC|LDCU|S2R|S2UR|R2UR|CS2R|ULDC)"), ("local (spill)", r"^(LDL|STL)"),
]
kernels = collections.OrderedDict()
seen = {}
name = None
for row in csv.reader(open(a.src)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        name = row[1]
        kernels.setdefault(name, [])
        continue
    if row[0] == "Address" or name is None:
        continue
    try:
        rec = (row[0], row[1].strip(), int(row[5]), int(row[2]))
    except (ValueError, IndexError):
        continue
    if rec[0] in seen.setdefault(name, set()):  # some ncu versions list every SASS line twice (SASS + source-correlated view)
        continue
    seen[name].add(rec[0])
    kernels[name].append(rec[1:])
for name, rows in kernels.items():
    if a.kernel and a.kernel not in name:
        continue
    total = sum(r[1] for r in rows)
    print(f"## {name[:140]}\n")
    print(f"SASS lines {len(rows)}, executed warp instructions {total:.4e}" + (f", {total * 32 / a.butterflies:.2f} thread instructions per butterfly" if a.butterflies else ""))
    by = collections.Counter()
    for text, n, _ in rows:
        op = re.sub(r"^@!?U?P\d+\s+", "", text)
        for label, rx in CLASSES:
            if re.match(rx, op):
                by[label] += n
                break
        else:
            by["other: " + op.split()[0]] += n
    print("\n| class | warp instr | share |" + (" per butterfly |" if a.butterflies else "") + "\n|---|---|---|" + ("---|" if a.butterflies else ""))
    for label, n in by.most_common(a.top):
        line = f"| {label} | {n:.4e} | {100 * n / total:.1f} % |"
        if a.butterflies:
            line += f" {n * 32 / a.butterflies:.2f} |"
        print(line)
    if a.stalls:
        print("\nmost-sampled instructions (stall samples):\n")
        tot_s = sum(r[2] for r in rows) or 1
        for text, n, s in sorted(rows, key=lambda r: -r[2])[:a.top]:
            print(f"    {100 * s / tot_s:5.1f} %  {text}")
    print()
