timeout 900 python -m pytest tests -m gpu -x -q -k "ik or shared_beta or task or config or body" 2>&1 | tail -3
timeout 300 python scripts/bench_ik_quick.py 2>&1 | tail -2
bash scripts/launch_list_ik.sh 2>&1 | tail -5
