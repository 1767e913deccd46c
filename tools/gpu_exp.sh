#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps', round(d['value']), 'e2e', round(d['e2e']['value']), d['extra'].get('single_frame_640x480_ms'))"
