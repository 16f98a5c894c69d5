"""SASS evidence for profiles/: per-kernel instruction counts (TMA tensor loads, mbarrier, cp.async, FMA flavours) of
libof3d.so and the TMA / mbarrier excerpt of the window z march.   python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections, re, subprocess, sys
LIB = 'opticalflow3d_dev_b200/libof3d.so'
sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in sass.split('\n'):
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur and re.match(r'\s+/\*[0-9a-f]{4,5}\*/', line):
        funcs[cur].append(line)
names = subprocess.run(['c++filt'], input='\n'.join(funcs), capture_output=True, text=True).stdout.split('\n')
KEEP = ('march_tz<of3d::TzSrcRaw<unsigned short', 'strip_conv2', 'march_window', 'strip_window_solve')
print('# SASS evidence, libof3d.so (sm_100a), round 2 final -- `python tools/sass_evidence.py` (cuobjdump -sass)')
print('# columns: instructions | UTMALDG (TMA tensor loads) | SYNCS (mbarrier) | LDGSTS (cp.async) | DFMA | FFMA | FFMA2 | kernel')
print()
rows = []
for (mangled, lines), dem in zip(funcs.items(), names):
    dem = dem.replace('of3d::', '')
    if not any(k.replace('of3d::', '') in dem for k in KEEP):
        continue
    cnt = lambda pat: sum(1 for l in lines if re.search(pat, l))
    short = re.sub(r'\(.*', '', dem)
    rows.append((short, '%6d | %3d | %3d | %4d | %5d | %5d | %5d | %s' % (len(lines), cnt(r'UTMALDG'), cnt(r'SYNCS'), cnt(r'LDGSTS'), cnt(r'\bDFMA'),
                                                                         cnt(r'\bFFMA\b'), cnt(r'FFMA2'), short)))
for _, r in sorted(rows):
    print(r)
for want in ('march_window_tma_sh<double, 25, 1, 2>', 'march_window_tma_sh<double, 49, 1, 1>'):
    for (mangled, lines), dem in zip(funcs.items(), names):
        if want in dem.replace('of3d::', '') and 'producer' not in dem:
            print('\n## excerpt: %s  (mbarrier protocol of the marching warps; the shifting ring: DFMA writes acc[i] from acc[i+1])' % want)
            shown = 0
            for l in lines:
                if re.search(r'SYNCS|UTMALDG', l) or (re.search(r'\bDFMA', l) and shown < 12):
                    print(l.rstrip()[:110]); shown += 1 if 'DFMA' in l else 0
    for (mangled, lines), dem in zip(funcs.items(), names):
        d = dem.replace('of3d::', '')
        if 'window_tma_producer<double, %s' % want.split('<double, ')[1].split(',')[0] in d:
            print('\n## excerpt: %s  (TMA tile loads of the producer warp)' % re.sub(r'\(.*', '', d))
            for l in lines:
                if re.search(r'SYNCS|UTMALDG', l):
                    print(l.rstrip()[:110])
            break
