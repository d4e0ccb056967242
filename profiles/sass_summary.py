"""SASS evidence per kernel: python profiles/sass_summary.py [lib] > profiles/rNN_sass_summary.txt
Counts the mnemonics that show tcgen05 / TMEM / bulk-async copies / mbarriers / cluster barriers / global atomics in every kernel of
the library and prints the MMA issue loop of the CTA-pair convolution."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "alphasnake_zero_b200/libasz_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
pat = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|LDTM(?:\.x\d+)?|UBLKCP\.S\.G|UBLKCP\.G\.S|UTCATOMSWS\S*|SYNCS\.\S+|"
                 r"UCGABAR_\w+|ELECT|ATOMG\.\S+|REDG?\.\S+|STS(?:\.\d+)?|LDS(?:\.\d+)?)\b")
fn, per, body = None, collections.OrderedDict(), collections.OrderedDict()
for l in sass:
    m = re.search(r"Function : (\S+)", l)
    if m:
        fn = m.group(1); per[fn] = collections.Counter(); body[fn] = []
        continue
    if fn and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l):
        per[fn]["_n"] += 1
        body[fn].append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip())
        for k in pat.findall(l):
            per[fn][k.split(".x")[0] if k.startswith("LDTM") else k] += 1
print("SASS evidence (cuobjdump -sass %s, CUDA 12.9, sm_100a; profiles/sass_summary.py).\n"
      "UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTCBAR = tcgen05.commit (.MULTICAST = to both CTAs of the pair), LDTM = tcgen05.ld,\n"
      "UTCATOMSWS = TMEM allocation, UBLKCP.S.G / .G.S = cp.async.bulk global->shared / shared->global, SYNCS.* = mbarrier operations,\n"
      "UCGABAR_* = cluster barrier, ATOMG / RED = global atomics.\n" % lib)
for fn, c in per.items():
    if not any(k.startswith(("UTC", "LDTM", "UBLK")) for k in c):
        continue
    dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    print("%s\n    %d SASS instructions; %s" % (dem[:160], c["_n"], ", ".join("%s x%d" % (k, c[k]) for k in sorted(c) if k != "_n")))
key = next((f for f in body if "conv_umma_kernelILi24ELb1ELb0" in f), None)
if key:
    b = body[key]
    idx = [i for i, l in enumerate(b) if "UTCHMMA" in l]
    print("\n---- excerpt: MMA issue loop of conv_umma_kernel<24, true, false> (one tap = 8 x UTCHMMA.2CTA + UTCBAR, then the next tap's barrier wait) ----")
    print("\n".join(b[max(0, idx[0] - 12):idx[min(17, len(idx) - 1)] + 6]))
key = next((f for f in body if "conv_umma_kernelILi24ELb1ELb1" in f), None)
if key:
    b = body[key]
    idx = [i for i, l in enumerate(b) if "UTCHMMA" in l]
    print("\n---- excerpt: first convolution, conv_umma_kernel<24, true, true>: five UTCHMMA.2CTA per 128 rows, two taps each (LBO in the A descriptor) ----")
    print("\n".join(b[max(0, idx[0] - 10):idx[-1] + 4]))
