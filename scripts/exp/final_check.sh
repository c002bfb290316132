#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --workload csr_ovo 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms', d['ms_per_step'], 'e2e', d['e2e'], 'df_s', d.get('dataframe_s'), 'cpu', d['cpu_baseline'])
"
