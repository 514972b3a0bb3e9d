#!/usr/bin/env python
"""tools/make_profiles.py REPORT.ncu-rep LAUNCHES.csv -- regenerate profiles/r1_qpd_full.md, r1_qpd_hotloop.md, r1_launches.{csv,md}
from one `ncu --set full --import-source on -k regex:k_qpd` report and one launch list of the bench command."""
import csv, io, os, shutil, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, launches = sys.argv[1], sys.argv[2]
P = os.path.join(ROOT, "profiles")
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, os.path.join(P, "r1_qpd_full.md")], check=True)
txt = open(os.path.join(P, "r1_qpd_full.md")).read()
open(os.path.join(P, "r1_qpd_full.md"), "w").write(
    "# r1 FINAL kernels: ncu --set full --clock-control none, `python tools/profile_step.py 1024 1` (config 2, one 1024-scenario step), kernels k_qpd<8|10|12>\n"
    "# (the first capture of the round, before the optimisations, is r1_qpd_full_v1.md)\n\n" + txt)
shutil.copy(launches, os.path.join(P, "r1_launches.csv"))
rows = list(csv.reader(open(launches)))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[h]
agg = defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    k = (d["Kernel Name"][:70], d["Grid Size"], d["Block Size"])
    agg[k][0] += 1
    agg[k][1] += float(d["Metric Value"].replace(",", ""))
tot = sum(v[1] for v in agg.values())
out = ["# r1 launch list (ncu --metrics gpu__time_duration.sum --clock-control none -c 400, `python bench.py --steps 2 --warmup 1 --streams 1`), final kernels of the round", "",
       "Per-launch times are cold-cache and serialised by ncu; only the SHARE of each kernel is meaningful. Source: `profiles/r1_launches.csv`.", "",
       "| kernel | grid | block | launches | total ms | ms / launch | share |", "|---|---|---|---|---|---|---|"]
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append("| `%s` | %s | %s | %d | %.3f | %.4f | %.1f %% |" % (k[0], k[1], k[2], c, t / 1e6, t / c / 1e6, 100 * t / tot))
open(os.path.join(P, "r1_launches.md"), "w").write("\n".join(out) + "\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2]
data = data[:len(data) // 2]
I = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0  # noqa: E731
isrc, ismp, iex, iw, iwi = (hdr.index(k) for k in ("Source", "# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"))
tot, ts, tw = (sum(I(r[c]) for r in data) for c in (iex, ismp, iw))
hot = [i for i, r in enumerate(data) if I(r[iex]) > 5_000_000]
h0 = h1 = hot[0]
for i in hot:
    if i - h1 <= 3:
        h1 = i
st = [hdr.index(k) for k in ("stall_barrier", "stall_short_sb", "stall_wait", "stall_long_sb", "stall_branch_resolving", "stall_math", "stall_not_selected", "stall_mio")]
rng = range(h0, h1 + 1)
lines = ["# r1 FINAL k_qpd<8> hot loop (qpd_block1, full-row layout: one ADMM iteration = S2 gather | S3 G.g | S1 row updates, 3 CTA barriers): ncu source page, per SASS instruction", "",
         "Capture: `ncu --set full --clock-control none --import-source on -k regex:k_qpd -c 3 python tools/profile_step.py 1024 1`, launch 0 (final kernel of the round).",
         "The loop holds %.0f %% of the kernel's executed warp-instructions, %.0f %% of its stall samples and %.0f %% of its shared-memory wavefronts (the rest: termination check every 25 iterations, block entry/exit, setup, polish)." % (
             100 * sum(I(data[i][iex]) for i in rng) / tot, 100 * sum(I(data[i][ismp]) for i in rng) / ts, 100 * sum(I(data[i][iw]) for i in rng) / tw),
         "Columns: row, SASS, warp-instructions executed (M), stall samples, shared wavefronts actual/ideal, samples by reason: barrier short_scoreboard wait long_scoreboard branch_resolving math_pipe not_selected mio.",
         "The v1 loop of the start of the round (loads consumed one by one, conflicting row numbering) is r1_qpd_hotloop_v1.md.", "", "```"]
for i in rng:
    r = data[i]
    lines.append("%d %s ex %5.1fM smp %6s wf %s/%s | %s" % (i, r[isrc][:48].ljust(48), I(r[iex]) / 1e6, r[ismp], r[iw], r[iwi], " ".join("%5s" % r[k] for k in st)))
lines.append("```")
open(os.path.join(P, "r1_qpd_hotloop.md"), "w").write("\n".join(lines) + "\n")
print(lines[3])
