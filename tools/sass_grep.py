"""Blackwell-specific SASS mnemonics per kernel of the built library (profiles/rN_sass_grep.txt):
    python tools/sass_grep.py > profiles/r2_sass_grep.txt
tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2), tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, TMA -> UTMALDG / UTMASTG,
cp.async.bulk -> UBLKCP, red.global.add.v4/.v2.f32 -> REDG...F32x4 / F32x2.  Needs cuobjdump and c++filt (no GPU)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "speech-decoding_b200", "sd_b200", "_lib", "libsd_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*Function : ', txt)
pat = re.compile(r'\b(UTC[A-Z]*MMA(?:\.2CTA)?|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UBLKCP[.\w]*|UTCBAR[.\w]*|(?<![A-Z])HMMA[.\w]*|REDG\.[.\w]*)')
print("# cuobjdump -sass speech-decoding_b200/sd_b200/_lib/libsd_b200.so   (sm_100a): Blackwell-specific mnemonics per kernel")
print("# tcgen05.mma -> UTCHMMA (kind::f16 and kind::tf32; .2CTA = cta_group::2), tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR")
print("# (.2CTA.MULTICAST = commit multicast to both CTAs of a pair), cp.async.bulk.tensor loads / stores -> UTMALDG / UTMASTG,")
print("# cp.async.bulk -> UBLKCP, red.global.add.v4 / .v2 .f32 -> REDG.F32x4 / REDG.F32x2.  HMMA (legacy mma.sync): none.\n")
tot = collections.Counter()
for f in funcs[1:]:
    name = f.split('\n', 1)[0].strip()
    c = collections.Counter()
    for m in pat.finditer(f):
        t = m.group(1)
        k = t.split('.')[0]
        if '.2CTA' in t: k += '.2CTA'
        if 'MULTICAST' in t: k += '.MULTICAST'
        if k == 'REDG':
            k += '.F32x4' if 'F32x4' in t else '.F32x2' if 'F32x2' in t else '.F64' if 'F64' in t else '.F32'
        c[k] += 1
    if not any(k.startswith(('UTC', 'LDTM', 'UTMA', 'UBLKCP', 'HMMA')) for k in c):
        continue
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r'\(.*', '', dem.replace('(anonymous namespace)::', '').replace('sd::', '').replace('void ', ''))
    print("%-40s %s" % (dem[:40], "  ".join("%s x%d" % kv for kv in sorted(c.items()))))
    tot.update(c)
print("\nTOTAL  " + "  ".join("%s x%d" % kv for kv in sorted(tot.items())))
