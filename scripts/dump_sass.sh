#!/bin/bash
# Regenerates profiles/sass/ from the library as built at head: one SASS listing per kernel (for the template families with many tuning
# variants — trace_kernel, trace_wide4_kernel, trace_mr_kernel — only the variants the defaults select), an instruction-mix summary
# (load widths / FMNMX3 / PRMT / vote instructions the designs rely on) and MANIFEST.json: EVERY kernel symbol of the .so with the listing
# that holds it (or the reason it has none).  tests/test_abi.py checks the manifest against the symbols of the built library, so a kernel
# added, removed or re-templated without re-running this script fails the CPU suite.
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import collections, json, os, re, shutil, subprocess, sys
sys.path.insert(0, os.getcwd())
from ntrace_b200 import build
LIB = 'ntrace_b200/libntrace_b200.so'
OUT = 'profiles/sass'
shutil.rmtree(OUT, ignore_errors=True)
os.makedirs(OUT)
txt = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
parts = re.split(r'\n\s*Function : ', txt)[1:]
names = subprocess.run(['c++filt'], input='\n'.join(p.split('\n', 1)[0] for p in parts), capture_output=True, text=True).stdout.split('\n')
# default variants of the tunable families: <layout, 128, smem 8, persistent, fast 0, wide 1> etc.
KEEP = {
    'trace_kernel': lambda a: a.split(', ')[1:3] == ['128', '8'] and a.split(', ')[5] in ('1', 'true') and a.split(', ')[6] in ('0', 'false'),
    'trace_wide4_kernel': lambda a: a.startswith('128, 8,') and a.split(', ')[-1] in ('1', 'true'),
    'trace_mr_kernel': lambda a: a.startswith('128, 8,'),
    'trace_sw_kernel': lambda a: a.startswith('128, 8,'),
}
summary, manifest, seen = [], [], collections.Counter()
for p, dem in zip(parts, names):
    dem = dem.strip()
    m = re.search(r'(\w+)<(.*)>\(', dem) or re.search(r'(\w+)\(', dem)
    base = m.group(1)
    targs = m.group(2) if m.lastindex and m.lastindex > 1 else ''
    targs_n = re.sub(r'\((?:int|bool|unsigned int)\)', '', targs)
    entry = {'kernel': base + (f'<{targs_n}>' if targs_n else '')}
    if base in KEEP and not KEEP[base](targs_n):
        entry['listing'] = None
        entry['why'] = 'tuning variant (NT_TRACE_* / NT_WIDE_* / NT_MR_* / NT_SW_* experiment knobs); the default variant of the family is listed'
        manifest.append(entry)
        continue
    tag = base + ('_' + re.sub(r'[^0-9a-zA-Z]+', '_', targs_n).strip('_') if targs_n else '')
    seen[tag] += 1
    if seen[tag] > 1:
        tag += f'_{seen[tag]}'
    body = p.split('\n', 1)[1]
    open(f'{OUT}/{tag}.sass', 'w').write('// ' + dem + '\n' + body)
    entry['listing'] = f'{tag}.sass'
    manifest.append(entry)
    ops = collections.Counter(re.findall(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', body, re.M))
    keys = ['LDG.E.256', 'LDG.E.128', 'LDG.E.CONSTANT', 'LDG.E.64', 'LDG.E', 'STG.E.128', 'FMNMX3', 'FMNMX', 'PRMT', 'VIMNMX', 'VOTE', 'VOTEU', 'SHFL', 'ATOMG', 'RED', 'LDS', 'STS', 'LDL', 'STL', 'MUFU.RCP', 'FFMA', 'FMUL', 'FADD', 'BAR', 'WARPSYNC', 'MATCH']
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
with open(f'{OUT}/SUMMARY.md', 'w') as f:
    f.write('# SASS instruction mix (cuobjdump -sass ntrace_b200/libntrace_b200.so; scripts/dump_sass.sh)\n\n')
    f.write(f'Library sources: sha256[:16] = `{build.source_sha16()}` (ntrace_b200.build.source_sha16).  ')
    f.write('Counts are static instructions by mnemonic prefix (`LDG.E` = all global loads; `.256` / `.128` / `.CONSTANT` count the loads carrying that qualifier).\n\n| kernel | instrs | mix |\n|---|---|---|\n')
    for tag, cnt, mix in summary:
        f.write(f'| `{tag}` | {cnt} | ' + ', '.join(f'{k} {v}' for k, v in mix.items()) + ' |\n')
json.dump({'source_sha16': build.source_sha16(), 'kernels': sorted(manifest, key=lambda e: e['kernel'])}, open(f'{OUT}/MANIFEST.json', 'w'), indent=1)
print(len(summary), 'listings,', len(manifest), 'kernel symbols')
PY
