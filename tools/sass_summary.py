"""Opcode / memory-instruction summary of the hot kernels from the built library
(cuobjdump -sass), written to profiles/r02_sass_summary.txt.  python tools/sass_summary.py"""
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, 'p1afempy_b200', 'libp1b200.so')
WANTED = [
    ('scatter (one GPU, pairs of triangles, 32-byte records)',
     r'scatter_kernel<p1::StiffnessValues, false, true, 0>'),
    ('scatter, mass matrix (16-byte slots {next, prev|k, v})',
     r'scatter_kernel<p1::MassValues, false, true, 4>'),
    ('scatter, slim slots (multi-GPU stiffness / mass)',
     r'scatter_kernel<p1::NoValues, false, false, 3>'),
    ('scatter, generalised stiffness (quadrature sums inside; two blocks per SM stated)',
     r'scatter_kernel_wide<p1::GeneralValues, false, false, 0>'),
    ('fill_rows (thread per row, 32-byte records)', r'fill_rows_kernel<false>'),
    ('fill_rows, mass matrix (16-byte slots)', r'fill_rows_kernel<true>'),
    ('fill_gather (multi-GPU: gathers coordinates, recomputes the entries)',
     r'fill_gather_kernel<p1::StiffnessValues>'),
    ('rowptr, tile sums', r'rowptr_kernel<0>'),
    ('rowptr, final pass', r'rowptr_kernel<1>'),
    ('rowptr, fused (one cooperative launch, small meshes)', r'rowptr_kernel<2>'),
    ('cold_init (counters, tile sums, status words)', r'cold_init_kernel'),
]
MEM = re.compile(r'^(LDG|STG|LD|ST|LDS|STS|LDL|STL|ATOMG|ATOMS|ATOM|REDG|RED|REDUX|LDSM|UTMA|UBLKCP)\b')
FP64 = re.compile(r'^(DADD|DMUL|DFMA|DSETP|MUFU)')
TENSOR = re.compile(r'^(HMMA|IMMA|DMMA|BMMA|UTC|TCGEN|UTMA|UBLKCP|LDGSTS)')


def main():
    names = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    blocks = {}
    cur = None
    for line in names.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            blocks[cur] = []
        elif cur and re.match(r'\s+/\*[0-9a-f]{4,}\*/', line):
            blocks[cur].append(line)
    demangled = subprocess.run(['c++filt'], input='\n'.join(blocks), capture_output=True,
                               text=True).stdout.splitlines()
    table = dict(zip(demangled, blocks.values()))
    out = ['# SASS summary of the hot kernels (round 2, final binary) -- cuobjdump -sass '
           'p1afempy_b200/libp1b200.so, sm_100a',
           '# What to look for: 256-bit global accesses (LDG/STG.E.ENL2.256: one 32-byte record per',
           '# instruction), the returning 64-bit atomic of the scatter (ATOMG.E.ADD.64), no tensor-core',
           '# or TMA instruction (the path is no contraction), fp64 division sequences (MUFU.RCP64H +',
           '# DFMA refinement) and -- compiled with -fmad=false -- DFMA only inside those sequences;',
           '# no local-memory traffic (LDL / STL) outside the rarely taken open-fan branch.', '']
    for title, pat in WANTED:
        hits = [k for k in table if pat in k]
        if not hits:
            out += ['## ' + title, 'NOT FOUND: ' + pat, '']
            continue
        name = hits[0]
        ops, mem, fp = collections.Counter(), collections.Counter(), collections.Counter()
        tensor = collections.Counter()
        for line in table[name]:
            body = re.sub(r'/\*.*?\*/', '', line).strip().rstrip(';').strip()
            body = re.sub(r'^@!?U?P\d+\s+', '', body)
            if not body:
                continue
            op = body.split()[0]
            base = op.split('.')[0]
            ops[base] += 1
            if MEM.match(op):
                mem[op] += 1
            if FP64.match(op):
                fp[op] += 1
            if TENSOR.match(op):
                tensor[op] += 1
        out += ['## ' + title, name, '%d SASS instructions' % sum(ops.values()),
                'by opcode: ' + ', '.join('%s %d' % kv for kv in ops.most_common(22)),
                'memory instructions: ' + ', '.join('%s x%d' % kv for kv in sorted(mem.items())),
                'fp64: ' + ', '.join('%s x%d' % kv for kv in sorted(fp.items())),
                'tensor-core / TMA instructions: ' +
                (', '.join('%s x%d' % kv for kv in tensor.items()) if tensor
                 else 'none (the path is no contraction)'), '']
    path = os.path.join(REPO, 'profiles', 'r02_sass_summary.txt')
    with open(path, 'w') as fh:
        fh.write('\n'.join(out))
    print(path)


if __name__ == '__main__':
    main()
