"""Static SASS opcode histogram of libogn.so, per kernel and in total (cuobjdump -sass; runs without a GPU).
usage: python tools/sass_histogram.py [out.json]     -> the evidence file under profiles/ (r02b_sass_opcodes.json)"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'origin_b200', 'lib', 'libogn.so')
# opcodes worth a column: TMA / bulk copies, mbarriers, FP32 / packed FP32 / FP64 math, min-max, tensor-core and TMEM
WATCH = ['UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'FFMA', 'FFMA2', 'FADD', 'FADD2', 'FMUL', 'FMUL2', 'FMNMX', 'FMNMX3', 'DFMA',
         'DMUL', 'DADD', 'LDCU', 'LDS', 'LDG', 'STG', 'STS', 'LDGSTS', 'R2UR', 'SHFL', 'ATOMG', 'RED', 'UTCHMMA', 'UTCQMMA',
         'UTCIMMA', 'LDTM', 'STTM', 'HMMA', 'DMMA']


def main(out_path):
    text = subprocess.run(['cuobjdump', '-sass', LIB], check=True, capture_output=True, text=True).stdout
    demangle = {}
    kernels, total = [], collections.Counter()
    name, counts = None, None
    for line in text.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            if name is not None:
                kernels.append((name, counts))
            name, counts = m.group(1), collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and counts is not None:
            op = m.group(1)
            counts['instructions'] += 1
            counts[op] += 1
    if name is not None:
        kernels.append((name, counts))
    names = [k for k, _ in kernels]
    if names:
        dem = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, dem))
    rows = []
    for k, c in kernels:
        total.update({op: n for op, n in c.items()})
        pretty = re.sub(r'\(.*', '', demangle.get(k, k))
        rows.append([pretty, {'instructions': c['instructions'], **{op: c[op] for op in WATCH if c[op]}}])
    doc = dict(library='origin_b200/lib/libogn.so (sm_100a), cuobjdump -sass, static instruction counts per kernel '
                       '(tools/sass_histogram.py)',
               total={op: total[op] for op in ['instructions'] + WATCH}, kernels=rows)
    with open(out_path, 'w') as f:
        json.dump(doc, f, indent=1)
    print({op: n for op, n in doc['total'].items() if n or op in ('UTMASTG', 'UTCHMMA', 'LDTM')})
    print(len(rows), 'kernels ->', out_path)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r02b_sass_opcodes.json'))
