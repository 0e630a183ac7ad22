timeout 240 python -m pytest tests/test_ae.py -m gpu -x -q 2>&1 | tail -30
echo "exit: $?"
