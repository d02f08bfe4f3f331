"""Per-instruction warp-stall samples of one kernel from an `ncu --set full --import-source on` report: totals per stall
reason over a SASS range, executed-instruction mix, and the hottest instructions. Used to split the attention kernel's
softmax loop into its phases (between the LDTM of the scores and the p_full arrive).

  python tools/ncu_stall_breakdown.py report.ncu-rep [--from LDTM.x32] [--to SYNCS.ARRIVE] [--top 20]
"""
import argparse
import collections
import csv
import io
import subprocess

NCU = "/usr/local/cuda/bin/ncu"
KEYS = ["stall_barrier", "stall_long_sb", "stall_math", "stall_mio", "stall_not_selected", "stall_selected", "stall_short_sb",
        "stall_wait", "stall_branch_resolving", "stall_dispatch", "stall_no_inst", "stall_lg", "stall_sleep", "stall_membar"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--from", dest="lo", default="LDTM.x32", help="first SASS line containing this starts the range (40 lines of lead-in are included)")
    ap.add_argument("--to", dest="hi", default="SYNCS.ARRIVE", help="first SASS line containing this, at least 300 lines later, ends it")
    ap.add_argument("--top", type=int, default=20)
    a = ap.parse_args()
    raw = subprocess.run([NCU, "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    ci = {h: i for i, h in enumerate(rows[hdr])}
    body = rows[hdr + 1:]
    first = next(i for i, r in enumerate(body) if a.lo in r[ci["Source"]])
    lo = max(first - 40, 0)
    hi = next(i for i, r in enumerate(body) if a.hi in r[ci["Source"]] and i > first + 300) + 5
    agg, mix = collections.Counter(), collections.Counter()
    total = 0
    per_iter = int(body[first][ci["Instructions Executed"]]) or 1
    for r in body[lo:hi]:
        total += int(r[ci["# Samples"]])
        for k in KEYS:
            if k in ci:
                agg[k] += int(r[ci[k]])
        op = r[ci["Source"]].split()
        op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "?")
        mix[op.split(".")[0]] += int(r[ci["Instructions Executed"]])
    print(f"SASS lines {lo}..{hi}: {total} samples")
    print("stall reasons:", {k.replace("stall_", ""): v for k, v in agg.most_common() if v})
    print("executed warp instructions per iteration:", {k: round(v / per_iter, 1) for k, v in mix.most_common(16)}, "total", round(sum(mix.values()) / per_iter, 1))
    hot = sorted(range(lo, hi), key=lambda i: -int(body[i][ci["# Samples"]]))[:a.top]
    for i in sorted(hot):
        r = body[i]
        print(i, r[ci["Source"]].strip()[:64], r[ci["# Samples"]],
              {k.replace("stall_", ""): int(r[ci[k]]) for k in KEYS if k in ci and int(r[ci[k]]) > 200})


if __name__ == "__main__":
    main()
