set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
lscpu | grep -E "Model name|^CPU\(s\)|Socket|Thread" 
free -g | head -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
timeout 600 python tools/probe.py 2>&1 | tail -20
