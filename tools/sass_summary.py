"""SASS opcode summary per kernel of libfactk.so (evidence of which kernels are tcgen05 / TMEM / TMA and which are legacy
mma.sync or CUDA-core code).  Runs without a GPU:  python tools/sass_summary.py > profiles/r2_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'fact_clip_b200', 'libfactk.so')
OPS = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'SYNCS', 'HMMA', 'LDSM', 'LDGSTS', 'FFMA', 'MUFU', 'UCGABAR', 'STS', 'LDS']
NOTE = {'UTCHMMA': 'tcgen05.mma (kind::f16 / tf32)', 'UTCBAR': 'tcgen05.commit', 'LDTM': 'tcgen05.ld', 'STTM': 'tcgen05.st',
        'UTMALDG': 'TMA load (cp.async.bulk.tensor)', 'SYNCS': 'mbarrier', 'HMMA': 'mma.sync', 'LDSM': 'ldmatrix', 'LDGSTS': 'cp.async',
        'UCGABAR': 'cluster barrier'}


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)', line)
        if m and cur is not None:
            cur[m.group(1).split('.')[0]] += 1
    names = subprocess.run(['c++filt'] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print('# SASS opcode counts per kernel of `fact_clip_b200/libfactk.so` (sm_100a) -- `python tools/sass_summary.py`\n')
    print('Legend: ' + ', '.join(f'`{k}` = {v}' for k, v in NOTE.items()) + '\n')
    print('| kernel | class | ' + ' | '.join(OPS) + ' | total |')
    print('|---|---|' + '---|' * (len(OPS) + 1))
    rows = []
    for (mangled, c), name in zip(kernels.items(), names):
        short = re.sub(r'\(.*', '', name).replace('factk::', '')
        cls = 'tcgen05 + TMEM + TMA' if c['UTCHMMA'] else ('mma.sync' if c['HMMA'] else 'CUDA cores')
        rows.append((cls, short, c))
    for cls, short, c in sorted(rows, key=lambda r: (r[0] != 'tcgen05 + TMEM + TMA', r[0] != 'mma.sync', r[1])):
        print(f'| `{short}` | {cls} | ' + ' | '.join(str(c[o]) if c[o] else '' for o in OPS) + f' | {sum(c.values())} |')


if __name__ == '__main__':
    main()
