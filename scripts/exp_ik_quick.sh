timeout 300 python scripts/bench_ik_quick.py 2>&1 | tail -2
bash scripts/launch_list_ik_modes.sh 2>&1 | tail -40
