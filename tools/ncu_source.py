#!/usr/bin/env python
"""Per-instruction warp-stall samples of one kernel launch in an .ncu-rep (debug tool).

usage:
  python tools/ncu_source.py REP [--launch K] marks            # landmark instructions (LDTM/STTM/BAR/SYNCS/UTCHMMA...)
  python tools/ncu_source.py REP [--launch K] buckets [STEP]   # samples per STEP-byte (hex, default 100) code bucket
  python tools/ncu_source.py REP [--launch K] lines LO HI      # every instruction in [LO, HI) (hex offsets)
  python tools/ncu_source.py REP [--launch K] ops LO HI        # samples per opcode in [LO, HI)
  python tools/ncu_source.py REP [--launch K] regions A:B:name ...  # samples of named ranges
Offsets are relative to the first instruction of the kernel, as cuobjdump -sass prints them.
"""
import collections
import csv
import io
import subprocess
import sys


def load(rep, launch):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip",
                          str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    ix = {h: i for i, h in enumerate(hdr)}
    seen, data = set(), []
    for r in rows:
        if len(r) == len(hdr) and r[0].startswith("0x") and r[0] not in seen:
            seen.add(r[0])
            data.append(r)
    base = int(data[0][0], 16)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    recs = []
    for r in data:
        st = {h[6:]: int(r[ix[h]]) for h in stalls if int(r[ix[h]]) > 0}
        recs.append((int(r[0], 16) - base, r[ix["Source"]].strip(), int(r[ix["# Samples"]]),
                     int(r[ix["Instructions Executed"]]), st))
    return recs


def opcode(src):
    t = src.split()
    return t[1] if t[0].startswith("@") else t[0]


def fmt_st(st, n=3):
    return sorted(st.items(), key=lambda kv: -kv[1])[:n]


def main():
    args = sys.argv[1:]
    rep = args.pop(0)
    launch = 0
    if args[0] == "--launch":
        launch = int(args[1])
        args = args[2:]
    cmd = args.pop(0)
    recs = load(rep, launch)
    tot = sum(r[2] for r in recs)
    print(f"# {rep} launch {launch}: {len(recs)} instructions, {tot} samples")
    if cmd == "marks":
        keys = ("LDTM", "STTM", "BAR.", "UTCHMMA", "UTCBAR", "USETMAXREG", "SYNCS.ARRIVE", "TRYWAIT", "UTMALDG",
                "UBLKCP", "EXIT", "MUFU.RCP", "WARPSYNC", "UTCATOMSWS")
        for a, src, s, ex, st in recs:
            if any(k in src for k in keys):
                print(f"{a:6x} {s:6d} ex={ex:9d} {src[:70]}")
    elif cmd == "buckets":
        step = int(args[0], 16) if args else 0x100
        b = collections.Counter()
        bs = collections.defaultdict(collections.Counter)
        for a, src, s, ex, st in recs:
            b[a // step] += s
            bs[a // step].update(st)
        for k in sorted(b):
            if b[k]:
                print(f"{k * step:6x} {b[k]:7d} {100 * b[k] / tot:5.1f}%  {bs[k].most_common(3)}")
    elif cmd == "lines":
        lo, hi = int(args[0], 16), int(args[1], 16)
        for a, src, s, ex, st in recs:
            if lo <= a < hi:
                print(f"{a:6x} {s:6d} ex={ex:9d} {src[:64]:64s} {fmt_st(st)}")
    elif cmd == "ops":
        lo, hi = int(args[0], 16), int(args[1], 16)
        agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
        for a, src, s, ex, st in recs:
            if lo <= a < hi:
                e = agg[opcode(src)]
                e[0] += 1
                e[1] += s
                e[2].update(st)
        for op, (n, s, st) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"{op:28s} n={n:4d} samples={s:7d} {st.most_common(4)}")
    elif cmd == "regions":
        for spec in args:
            lo, hi, name = spec.split(":", 2)
            lo, hi = int(lo, 16), int(hi, 16)
            s = sum(r[2] for r in recs if lo <= r[0] < hi)
            st = collections.Counter()
            for r in recs:
                if lo <= r[0] < hi:
                    st.update(r[4])
            print(f"{name:40s} {s:7d} {100 * s / tot:5.1f}%  {st.most_common(4)}")
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
