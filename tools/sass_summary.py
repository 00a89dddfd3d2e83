"""Blackwell evidence: per kernel of the built library, how many tcgen05 / TMEM / TMA instructions its SASS holds
(cuobjdump -sass | count of UTC*MMA, LDTM, STTM, UBLKCP, UTMALDG, UTMASTG, LDGSTS, HMMA per function).
usage: python tools/sass_summary.py [lib.so] > profiles/r2_sass_tcgen05.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'instancerefer_b200', 'libinstancerefer_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
MN = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTCCP', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'LDGSTS', 'HMMA', 'SYNCS', 'UTCBAR']
fn, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        fn = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn:
        m = re.search(r'\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)', line)
        if m:
            op = m.group(1).split('.')[0]
            if op in MN:
                counts[fn][op] += 1
            if op == 'UTMALDG' and 'GATHER4' in m.group(1):
                counts[fn]['UTMALDG.GATHER4'] += 1
print(f'# cuobjdump -sass {os.path.basename(lib)}: instruction counts per kernel (only kernels with at least one of them)')
print('# UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk, UTMALDG = cp.async.bulk.tensor (TMA), LDGSTS = cp.async')
tot = collections.Counter()
for f, c in counts.items():
    if sum(c.values()) == 0:
        continue
    tot.update(c)
    print(f'{f[:110]:110s} ' + ' '.join(f'{k}={v}' for k, v in sorted(c.items())))
print('TOTAL ' + ' '.join(f'{k}={v}' for k, v in sorted(tot.items())))
