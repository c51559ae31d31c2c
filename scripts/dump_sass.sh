#!/bin/bash
# Regenerates profiles/sass/: one SASS listing per kernel of the built library (default trace variants only) and an
# instruction-mix summary that shows the load widths / FMNMX3 / vote instructions the design relies on.
set -e
cd "$(dirname "$0")/.."
LIB=ntrace_b200/libntrace_b200.so
OUT=profiles/sass
rm -rf $OUT && mkdir -p $OUT
cuobjdump -sass $LIB > /tmp/nt_all.sass
python - <<'PY'
import re, subprocess, collections, os
txt = open('/tmp/nt_all.sass').read()
parts = re.split(r'\n\s*Function : ', txt)[1:]
names = subprocess.run(['c++filt'], input='\n'.join(p.split('\n', 1)[0] for p in parts), capture_output=True, text=True).stdout.split('\n')
summary = []
seen = collections.Counter()
for p, dem in zip(parts, names):
    m = re.search(r'(\w+)<(.*?)>\(', dem) or re.search(r'(\w+)\(', dem)
    base = m.group(1)
    targs = m.group(2) if m.lastindex and m.lastindex > 1 else ''
    if base == 'trace_kernel' and targs not in ('4, 128, 8, true, 0, false, true', '4, 128, 8, false, 0, false, true', '5, 128, 8, true, 0, false, true'):
        continue          # tuning variants (NT_TRACE_* env knobs); the three kept are the defaults: Compact persistent / non-persistent, Compact2
    tag = base + ('_' + re.sub(r'[^0-9a-zA-Z]+', '_', targs).strip('_') if targs else '')
    seen[tag] += 1
    if seen[tag] > 1:
        tag += f'_{seen[tag]}'
    body = p.split('\n', 1)[1]
    open(f'profiles/sass/{tag}.sass', 'w').write('// ' + dem + '\n' + body)
    ops = collections.Counter(re.findall(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', body, re.M))
    keys = ['LDG.E.256', 'LDG.E.128', 'LDG.E.CONSTANT', 'LDG.E.64', 'LDG.E', 'STG.E.128', 'FMNMX3', 'FMNMX', 'VOTE', 'VOTEU', 'SHFL', 'ATOMG', 'RED', 'LDS', 'STS', 'LDL', 'STL', 'MUFU.RCP', 'FFMA', 'FMUL', 'FADD', 'BAR', 'WARPSYNC', 'MATCH']
    def n(k):
        if k == 'LDG.E.256':
            return sum(v for o, v in ops.items() if o.startswith('LDG') and '.256' in o)
        if k == 'LDG.E.128':
            return sum(v for o, v in ops.items() if o.startswith('LDG') and '.128' in o)
        if k == 'LDG.E.CONSTANT':
            return sum(v for o, v in ops.items() if o.startswith('LDG') and 'CONSTANT' in o)
        return sum(v for o, v in ops.items() if o.startswith(k))
    mix = {k: n(k) for k in keys}
    summary.append((tag, sum(ops.values()), {k: v for k, v in mix.items() if v}))
with open('profiles/sass/SUMMARY.md', 'w') as f:
    f.write('# SASS instruction mix (cuobjdump -sass ntrace_b200/libntrace_b200.so; scripts/dump_sass.sh)\n\n')
    f.write('Counts are static instructions by mnemonic prefix (`LDG.E` = all global loads; `.256` / `.128` / `.CONSTANT` count the loads carrying that qualifier).\n\n| kernel | instrs | mix |\n|---|---|---|\n')
    for tag, n, mix in summary:
        f.write(f'| `{tag}` | {n} | ' + ', '.join(f'{k} {v}' for k, v in mix.items()) + ' |\n')
print(len(summary), 'kernels')
PY
