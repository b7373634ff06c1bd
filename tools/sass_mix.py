"""Static SASS opcode mix of named kernels of the built library (no GPU needed):
    python tools/sass_mix.py "ks_accumulate_kernel<4, true, true>" "ntt_inv_fp_kernel<13>"
Counts every instruction of the kernel text (all branch arms, unrolled bodies): not issued-instruction counts."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "prefhetch_b200" / "libprefhetch_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
parts = re.split(r'\n\s*Function : ', txt)
want = sys.argv[1:]
def demangle(n):
    return subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
for p in parts[1:]:
    name = p.split('\n', 1)[0].strip()
    dn = demangle(name)
    short = dn.split('(')[0].replace('void ', '')
    if short not in want:
        continue
    ops = collections.Counter()
    n = 0
    for line in p.split('\n'):
        m = re.search(r'/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if not m:
            continue
        op = m.group(1)
        n += 1
        base = op.split('.')[0]
        if base == 'IMAD':
            base = 'IMAD.WIDE' if '.WIDE' in op else ('IMAD.HI' if '.HI' in op else ('IMAD.MOV' if '.MOV' in op or '.SHL' in op or '.IADD' in op else 'IMAD'))
        if base in ('LDG', 'STG', 'LDS', 'STS', 'LD', 'ST'):
            w = re.search(r'\.(128|64|U8|U16|S8)', op)
            base += ('.' + w.group(1)) if w else '.32'
        ops[base] += 1
    print(f"== {short}: {n} SASS instructions (static)")
    print("   " + ", ".join(f"{k} {v}" for k, v in ops.most_common(22)))
