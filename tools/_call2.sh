set -x
python -m pytest tests/test_gpu_next_rows.py -x -q 2>&1 | tail -40
