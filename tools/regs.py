"""Summarise ptxas -v output of the marching kernels from opticalflow3d_dev_b200/build.log."""
import re, subprocess, sys
txt = open('opticalflow3d_dev_b200/build.log').read()
ents = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:.*\n)*?ptxas info\s+: Function properties for [^\n]+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt)
names = subprocess.run(['c++filt'], input='\n'.join(e[0] for e in ents), capture_output=True, text=True).stdout.split('\n')
filt = sys.argv[1] if len(sys.argv) > 1 else 'march'
for (name, stack, ss, sl, regs), dem in sorted(zip(ents, names), key=lambda x: x[1]):
    if filt not in dem: continue
    m = re.match(r'void of3d::(\w+)<([^>]*)>', dem)
    print('%-14s %-40s regs %3s stack %4s spill %5s/%5s' % (m.group(1), m.group(2), regs, stack, ss, sl))
