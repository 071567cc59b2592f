"""SASS opcode census of the in-tree library: which Blackwell instructions each kernel family really contains
(the PTX names never appear in SASS: tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG / UTMASTG,
cp.async.bulk -> UBLKCP, tcgen05.commit -> UTCBAR).   python tools/sass_census.py > profiles/rNN_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'reliability-challenges-uncertainty_b200', 'librcu_b200.so')
OPS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'HMMA', 'LDGSTS']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0].replace('void ', '')
            per.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r'^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m:
            per[cur][m.group(1).split('.')[0]] += 1
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print('# SASS opcode census of librcu_b200.so (sm_100a)\n')
    print('`cuobjdump -sass reliability-challenges-uncertainty_b200/librcu_b200.so`, instruction counts per kernel; `HMMA` (legacy mma.sync) must be 0.\n')
    print('| kernel | ' + ' | '.join(OPS) + ' | instructions |')
    print('|---|' + '---|' * (len(OPS) + 1))
    for name, c in per.items():
        if not any(c[o] for o in OPS[:6]) and 'conv' not in name:
            continue
        print('| `%s` | ' % name[:90] + ' | '.join(str(c[o]) for o in OPS) + ' | %d |' % sum(c.values()))
    print('| **whole library (%d kernels)** | ' % len(per) + ' | '.join('**%d**' % tot[o] for o in OPS) + ' | %d |' % sum(tot.values()))


if __name__ == '__main__':
    sys.exit(main())
