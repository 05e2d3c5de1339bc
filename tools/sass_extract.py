#!/usr/bin/env python
"""Counts the Blackwell-native SASS mnemonics per kernel of libuocr.so (cuobjdump -sass) -> profiles/rNN_sass_tcgen05.txt.

    python tools/sass_extract.py [profiles/r02_sass_tcgen05.txt]

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit,
SYNCS = mbarrier operations; HMMA (legacy mma.sync) must not appear.  Runs without a GPU.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'univer_ocr_b200', 'lib', 'libuocr.so')
PATTERNS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'SYNCS', 'HMMA', 'FMUL2', 'FFMA2']


def main():
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r02_sass_tcgen05.txt')
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    per_kernel, current = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            current = m.group(1)
            per_kernel[current] = collections.Counter()
            continue
        if current is None:
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
        if not m:
            continue
        op = m.group(1)
        per_kernel[current]['_total'] += 1
        base = op.split('.')[0]
        if base in PATTERNS:
            per_kernel[current][base] += 1
        if op.startswith('UTCHMMA') and '.2CTA' in op:
            per_kernel[current]['UTCHMMA.2CTA'] += 1
    demangle = subprocess.run(['c++filt'], input='\n'.join(per_kernel), capture_output=True, text=True).stdout.splitlines()
    totals = collections.Counter()
    with open(dst, 'w') as out:
        out.write('# cuobjdump -sass univer_ocr_b200/lib/libuocr.so: Blackwell-native mnemonics per kernel (tools/sass_extract.py)\n')
        out.write('# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA load, UTCBAR = tcgen05.commit, SYNCS = mbarrier\n')
        cols = [p for p in PATTERNS]
        out.write(f'{"instr":>7s} ' + ' '.join(f'{c:>12s}' for c in cols) + '  kernel\n')
        for (name, cnt), pretty in zip(per_kernel.items(), demangle):
            if not any(cnt[c] for c in cols if c not in ('SYNCS', 'FMUL2', 'FFMA2')):
                continue
            pretty = pretty.replace('(anonymous namespace)::', '')
            pretty = re.sub(r'\(.*', '', pretty).replace('void ', '').replace('uocr::', '')
            out.write(f'{cnt["_total"]:7d} ' + ' '.join(f'{cnt[c]:12d}' for c in cols) + f'  {pretty}\n')
            totals.update({c: cnt[c] for c in cols})
        out.write(f'{"total":>7s} ' + ' '.join(f'{totals[c]:12d}' for c in cols) + f'  ({len(per_kernel)} kernels in the library)\n')
    print(open(dst).read())


if __name__ == '__main__':
    main()
