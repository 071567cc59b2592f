"""Exhaustive check (authoring-time experiment, results recorded in DESIGN.md):

For the binary case the eval-side uncertainty (rechun/eval/analysis.py:201 -> numpyfunctions.py:166-168)
is a function of the single saved float32 foreground probability p:
    u(p) = -( where(q>0, q*log(q), 0.0) + where(p>0, p*log(p), 0.0) ) / ln 2,   q = fl32(1 - p)
Walk ALL float32 values in [0, 1] in increasing order and, for each sweep threshold, record every position
where the predicate u(p) > th flips.  If each threshold flips exactly twice (off->on, on->off) the set
{p : u(p) > th} is one float32 interval and a kernel may classify p by interval — bit-exactly.
"""
import sys
import numpy as np

THRESHOLDS = (0.05, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 0.95)


def u_of_p(p):
    prob = np.stack([1 - p, p], axis=-1)
    return -np.where(prob > 0, prob * np.log(prob), [0.0]).sum(axis=-1) / np.log(2)


def main():
    one_bits = int(np.float32(1.0).view(np.uint32))
    chunk = 1 << 23
    flips = {th: [] for th in THRESHOLDS}
    last = {th: False for th in THRESHOLDS}
    with np.errstate(divide='ignore', invalid='ignore'):
        for start in range(0, one_bits + 1, chunk):
            bits = np.arange(start, min(start + chunk, one_bits + 1), dtype=np.uint32)
            p = bits.view(np.float32)
            u = u_of_p(p)
            for th in THRESHOLDS:
                on = u > th
                d = np.flatnonzero(on[1:] != on[:-1])
                if on[0] != last[th]:
                    flips[th].append(int(bits[0]))
                flips[th].extend(int(bits[i + 1]) for i in d)
                last[th] = bool(on[-1])
            if (start // chunk) % 16 == 0:
                print('.. %5.1f%%' % (100.0 * start / one_bits), flush=True)
    ok = True
    for th in THRESHOLDS:
        f = flips[th]
        print(th, len(f), [hex(b) for b in f[:6]], [float(np.uint32(b).view(np.float32)) for b in f[:6]])
        ok &= len(f) == 2
    print('SINGLE_INTERVAL_FOR_ALL_THRESHOLDS', ok)
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
