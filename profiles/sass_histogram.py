#!/usr/bin/env python
"""Opcode histogram of the engine library's SASS (cuobjdump -sass): which Blackwell-native / tensor / async-copy
instructions the hand-written kernels compile to.   python profiles/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'occuspytial_b200', 'liboccu_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
full = collections.Counter()
for line in sass.splitlines():
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Za-z0-9_.]*)', line)
    if m:
        full[m.group(1)] += 1
base = collections.Counter()
for op, c in full.items():
    base[op.split('.')[0]] += c
total = sum(full.values())
print(f'# SASS opcode histogram of occuspytial_b200/liboccu_b200.so (sm_100a), round 2 (profiles/sass_histogram.py)')
print(f'# cuobjdump -sass | opcode field; {total} instructions in all kernels\n')
print('## Blackwell-native / tensor / async-copy opcodes')
for key in ('UTCIMMA', 'UTCBAR', 'UTMALDG', 'LDTM', 'DMMA', 'LDGSTS', 'SYNCS', 'MUFU', 'DFMA', 'DMUL', 'DADD', 'FFMA', 'IMAD', 'LOP3', 'LDG',
            'STG', 'LDS', 'STS', 'BAR', 'SHFL', 'VOTE', 'WARPSYNC'):
    variants = sorted(((c, op) for op, c in full.items() if op.split('.')[0] == key), reverse=True)
    if variants:
        print(f'{key:12s} {base[key]:6d}   ' + ', '.join(f'{op} x{c}' for c, op in variants[:12]))
print('\n## top 40 base opcodes')
for op, c in base.most_common(40):
    print(f'{op:12s} {c:6d}')
