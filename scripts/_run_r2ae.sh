mkdir -p gpurun_out
PROF_B=128 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'raycast_packed' -c 2 -o gpurun_out/prof_raycast -f python scripts/time_raycast.py > gpurun_out/ncu_raycast.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_raycast.ncu-rep --page details --csv 2>/dev/null | python - <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
seen = set()
for r in rows[1:]:
    if r[ix['ID']] != rows[1][ix['ID']]:
        continue
    name = r[ix['Metric Name']]
    if any(k in name for k in ('Duration', 'Issue Slots Busy', 'Executed Ipc', 'Pipe', 'Stall', 'Warp Cycles Per Issued', 'Achieved Occupancy', 'Theoretical Occupancy', 'Eligible', 'Issued Warp', 'No Eligible', 'L1/TEX Hit', 'Shared', 'Registers', 'Bank')):
        print('%-28s %-48s %s %s' % (r[ix['Section Name']][:28], name[:48], r[ix['Metric Value']], r[ix['Metric Unit']]))
PY
ncu -i gpurun_out/prof_raycast.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
r = rows[2] if len(rows) > 2 else None
if r:
    for h, v in zip(hdr, r):
        if ('pcsamp' in h or 'inst_executed_pipe' in h or 'issue_active' in h or 'warp_issue_stalled' in h) and v not in ('', '0'):
            print(h, v)
PY
