"""Decode the scheduling control fields (stall count, yield, write/read scoreboard, wait mask) of sm_100 SASS.

    cuobjdump -sass -fun <mangled> obj.o | python tools/sass_sched.py [--from ADDR] [--to ADDR] [--grep MUFU]

Prints one line per instruction with its static stall count, and a static issue-cycle estimate (sum of stall counts) between
consecutive MUFU.EX2 -- the tool the exponential loop of the attention kernels was tuned with (ptxas decides the SASS order;
the only way to see whether a MUFU's consumers sit far enough behind it is to read it).
"""
import re
import sys


def parse(text):
    lines = text.split("\n")
    recs = []
    i = 0
    while i < len(lines):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                recs.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=(hi >> 41) & 0xF, yld=(hi >> 45) & 1,
                                 wr=(hi >> 46) & 7, rd=(hi >> 49) & 7, wait=(hi >> 52) & 0x3F))
                i += 2
                continue
        i += 1
    return recs


def main():
    args = sys.argv[1:]
    lo, hi, summary, fun = 0, 1 << 62, False, None
    while args:
        a = args.pop(0)
        if a == "--from":
            lo = int(args.pop(0), 16)
        elif a == "--to":
            hi = int(args.pop(0), 16)
        elif a == "--summary":
            summary = True
        elif a == "--fun":          # substring of the (mangled) function name: only that function of a whole-file dump
            fun = args.pop(0)
    text = sys.stdin.read()
    if fun is not None:
        parts = re.split(r"(?m)^\s*Function : ", text)
        text = "\n".join(p for p in parts[1:] if fun in p.split("\n", 1)[0])
    recs = [r for r in parse(text) if lo <= r["addr"] <= hi]
    if summary:
        # distance (in static issue cycles and instructions) from each MUFU.EX2 to the first instruction that waits on its scoreboard
        cyc = 0
        pend = {}
        dists = []
        n_mufu = 0
        for r in recs:
            if r["wait"]:
                for sb in range(6):
                    if (r["wait"] >> sb) & 1 and sb in pend:
                        dists.append(cyc - pend.pop(sb))
            if "MUFU.EX2" in r["text"]:
                n_mufu += 1
                if r["wr"] != 7:
                    pend[r["wr"]] = cyc
            cyc += max(r["stall"], 1)
        print(f"instructions {len(recs)}, MUFU.EX2 {n_mufu}, static issue cycles {cyc} ({cyc / max(n_mufu, 1):.1f} per MUFU)")
        if dists:
            dists.sort()
            print(f"scoreboard wait distance (static cycles after the MUFU that set it): min {dists[0]}, median {dists[len(dists) // 2]}, max {dists[-1]}, n {len(dists)}")
        return
    for r in recs:
        print(f"{r['addr']:05x} st={r['stall']:2d} y={r['yld']} wr={r['wr']} rd={r['rd']} wait={r['wait']:06b}  {r['text'][:100]}")


if __name__ == "__main__":
    main()
