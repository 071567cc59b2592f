"""Barrier-wait and tensor-issue census of ONE kernel from an ncu report captured with --set full --import-source on:
every mbarrier wait site (first try + polling loop) with its executions and warp-stall samples, the tcgen05.mma / tcgen05.ld /
TMA instructions, spill traffic and shared-memory wavefronts by opcode.   python tools/ncu_waits.py rep.ncu-rep > profiles/x.md
"""
import collections
import csv
import subprocess
import sys


def main(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    name = rows[0][1] if rows and len(rows[0]) > 1 else '?'
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    total = sum(f(r, '# Samples') for r in data)
    print('kernel: `%s`\n\ntotal warp-stall samples: %d\n' % (name, total))
    print('| SASS index | instruction | executed | samples | share |\n|---|---|---|---|---|')
    groups = collections.OrderedDict()
    for i, r in enumerate(data):
        s = r[ix['Source']].strip()
        op = [t for t in s.split() if not t.startswith('@')]
        op = op[0] if op else ''
        key = None
        if 'TRYWAIT' in s:
            key = 'mbarrier try_wait'
        elif op.startswith('UTCHMMA'):
            key = 'UTCHMMA (tcgen05.mma)'
        elif op.startswith('LDTM'):
            key = 'LDTM (tcgen05.ld)'
        elif op.startswith('UTMALDG') or op.startswith('UTMASTG') or op.startswith('UTCBAR') or op.startswith('UBLKCP'):
            key = op.split('.')[0]
        elif op.startswith('BAR'):
            key = 'BAR.SYNC'
        elif op.startswith('LDL') or op.startswith('STL'):
            key = 'spill LDL / STL'
        if key is None:
            continue
        g = groups.setdefault(key, [0, 0.0, 0.0])
        g[0] += 1
        g[1] += f(r, 'Instructions Executed')
        # samples of a wait site sit on the branch that follows the try_wait
        smp = f(r, '# Samples')
        if 'TRYWAIT' in s:                       # ... up to and including the next branch
            for j in range(i + 1, min(i + 4, len(data))):
                smp += f(data[j], '# Samples')
                if 'BRA' in data[j][ix['Source']]:
                    break
        g[2] += smp
        if 'TRYWAIT' in s:
            print('| %d | `%s` | %d | %d | %.3f |' % (i, s[:60], f(r, 'Instructions Executed'), smp, smp / max(1.0, total)))
    print('\n| class | sites | executed | samples | share |\n|---|---|---|---|---|')
    for k, (n, e, smp) in groups.items():
        print('| %s | %d | %d | %d | %.3f |' % (k, n, e, smp, smp / max(1.0, total)))
    wf = collections.Counter()
    for r in data:
        s = r[ix['Source']].split()
        op = [t for t in s if not t.startswith('@')]
        if op:
            wf[op[0].split('.')[0]] += f(r, 'L1 Wavefronts Shared')
    print('\nshared-memory wavefronts by opcode (LSU side): ' + ', '.join('%s %.2f M' % (k, v / 1e6) for k, v in wf.most_common(4) if v > 0))


if __name__ == '__main__':
    main(sys.argv[1])
