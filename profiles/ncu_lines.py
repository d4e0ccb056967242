"""Summarise an ncu report per CUDA source line: python profiles/ncu_lines.py <rep> <kernel-regex> [top]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, out, seen_fn = None, None, [], 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ix = {}; [ix.setdefault(h, i) for i, h in enumerate(hdr)]; continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        def f(k):
            try: return float(r[ix[k]])
            except ValueError: return 0.0
        out.append((cur, r[0], r[1], f("Instructions Executed"), f("# Samples"), f("stall_long_sb"), f("stall_short_sb"), f("stall_wait"), f("stall_barrier"), f("stall_branch_resolving")))
tot = sum(o[3] for o in out) or 1; ts = sum(o[4] for o in out) or 1
print("total warp-instructions %.0f, samples %.0f" % (tot, ts))
out.sort(key=lambda o: -o[4])
print("%6s %6s | lsb ssb wait bar br | line" % ("inst%", "smp%"))
for o in out[:top]:
    print("%5.1f%% %5.1f%% | %3.0f %3.0f %3.0f %3.0f %3.0f | %s:%s %s" % (100 * o[3] / tot, 100 * o[4] / ts, 100*o[5]/ts, 100*o[6]/ts, 100*o[7]/ts, 100*o[8]/ts, 100*o[9]/ts, o[0], o[1], o[2].strip()[:95]))
