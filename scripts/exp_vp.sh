timeout 900 python -m pytest tests -m gpu -x -q -k "vposer or decoder or config" 2>&1 | tail -3
bash scripts/launch_list_ik_modes.sh 2>&1 | grep -i 'vposer\|contract' | tail -6
