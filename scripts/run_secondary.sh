./tests/cpp/test_host 2>&1 | tail -15
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12
