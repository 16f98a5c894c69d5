"""DRAM traffic per stage / per voxel from an `ncu --csv` launch list (gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum per launch) -> the JSON bench.py reads for `roofline.traffic` (profiles/r02_traffic.json).

    python tools/ncu_traffic.py key=launches.csv[:voxels_per_timepoint[:launches_to_skip[:launches_per_timepoint]]] ... > profiles/r02_traffic.json
"""
import collections
import csv
import json
import sys

STAGES = (('temporal', 'temporal'), ('march_tz', 'gradient_z'), ('strip_conv2', 'gradient_xy'), ('march_window', 'products_window_z'),
          ('strip_window_solve', 'window_xy_solve'))


def parse(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, mi, ii = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name'), hdr.index('ID')
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = float(r[vi].replace(',', ''))
    return d


out = {}
for arg in sys.argv[1:]:
    key, rest = arg.split('=', 1)
    parts = rest.split(':')
    path = parts[0]
    vox = int(parts[1]) if len(parts) > 1 else 128 * 1024 * 1024
    skip = int(parts[2]) if len(parts) > 2 else 0
    per_tp = int(parts[3]) if len(parts) > 3 else None
    launches = [(k, m) for (i, k), m in parse(path).items()][skip:]
    if per_tp:
        launches = launches[:per_tp]
    st_bytes, st_ms = collections.OrderedDict(), collections.OrderedDict()
    for k, m in launches:
        stage = next((s for pat, s in STAGES if pat in k), None)
        if stage is None:
            continue
        st_bytes[stage] = st_bytes.get(stage, 0.0) + m.get('dram__bytes_read.sum', 0.0) + m.get('dram__bytes_write.sum', 0.0)
        st_ms[stage] = st_ms.get(stage, 0.0) + m.get('gpu__time_duration.sum', 0.0) / 1e6
    total = sum(st_bytes.values())
    out[key] = {'bytes_per_voxel': total / vox, 'kernels': len(launches), 'source': path,
                'stage_dram_bytes_per_timepoint': st_bytes, 'stage_ms_per_timepoint_ncu_cold': st_ms,
                'stage_dram_bytes_per_launch': {s: st_bytes[s] for s in ('window_xy_solve', 'products_window_z', 'gradient_z') if s in st_bytes}}
print(json.dumps(out, indent=1))
