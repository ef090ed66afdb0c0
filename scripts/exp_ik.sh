timeout 600 python -m pytest tests/test_ik_gpu.py -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
from smplpp_b200 import ik_bench
dev = torch.device("cuda", 0)
def barrier(): torch.cuda.synchronize()
print(json.dumps(ik_bench.run(dev, 0, 1, lambda x: x, barrier)))
PY
