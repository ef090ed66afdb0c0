#!/bin/bash
timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ik | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['ms_per_launch'])"
