"""Per kernel of libgsplat_b200.so: instruction count and the SASS mnemonics that prove the asynchronous machinery
(UBLKCP = cp.async.bulk / 1-D TMA, LDGSTS = cp.async, SYNCS = mbarrier, REDG = red.global.add, MUFU = ex2 / rcp / rsqrt).
Usage: python scripts/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pf3plat_b200", "csrc", "libgsplat_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, c = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        c[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        c[fn][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(c.keys()), capture_output=True, text=True).stdout.splitlines()
keys = ["UBLKCP", "LDGSTS", "SYNCS", "REDG", "ATOMG", "MUFU", "VOTE", "SHFL", "BAR", "LDS", "STS", "FFMA"]
print(__doc__.strip().splitlines()[0])
for (k, cnt), nm in sorted(zip(c.items(), names), key=lambda t: t[1]):
    short = nm.replace("(anonymous namespace)::", "")
    short = re.sub(r"^void ", "", short)
    short = re.sub(r"\(.*", "", short)
    if short.startswith("cub::") or short.startswith("thrust::"):
        continue
    print(f"{short}: total={sum(cnt.values())} " + " ".join(f"{k2}={cnt[k2]}" for k2 in keys if cnt[k2]))
